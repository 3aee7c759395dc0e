"""The C++ host mirror of the reference API (godot-audio-spatializer_b200/host: AudioSpatializer3D,
AudioSpatializerInstance3D, SpatializerParameters3D, BatchMixer ...), driven by tests/host/host_test.cpp.

CPU: class defaults and setter validation (the reference's ERR_FAIL_* conditions).
GPU: a scene played through BatchMixer (built-in 3D instances + one custom instance that overrides
     calculate_spatialization) must match the oracle driven through the plain C-ABI flow on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import scenarios as S

abi, synth = S.abi, S.synth
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "host", "host_test")


@pytest.fixture(scope="module")
def host_test():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "godot-audio-spatializer_b200", "host")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host")])
    return BIN


def test_validation_matches_reference_rules(host_test):
    out = subprocess.run([host_test, "validate"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "validate ok" in out.stdout
    # the refused setters print like ERR_FAIL_* does
    assert "Panning strength must be a positive number." in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode_b", [0, 1])
def test_scene_through_batch_mixer_matches_oracle(host_test, orc, tmp_path, mode_b):
    V, F, blocks, speaker_mode, num_buses = 40, 256, 3, abi.SPEAKER_SURROUND_51, 2
    area = synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)
    emitters = [synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.5) for b in range(blocks)]
    sources = [synth.make_sources(V, F, block=b, mix_rate=48000.0) for b in range(blocks)]
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.array([V, F, blocks, speaker_mode, num_buses, mode_b, 1, 1], dtype=np.int32).tobytes())
        f.write(np.asarray(area, dtype=abi.area).tobytes())
        for b in range(blocks):
            f.write(emitters[b].tobytes())
            f.write(sources[b].tobytes())
    out = subprocess.run([host_test, "scene", str(inp), str(outp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    C = speaker_mode + 1
    got = np.fromfile(outp, dtype=np.float32).reshape(blocks, num_buses, C, F, 2)

    # the same scene through the C-ABI flow on the oracle: built-in instances via gain_compute, the custom last
    # instance via params_set with what FixedInstance::calculate_spatialization returns
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=4, num_buses=num_buses, speaker_mode=speaker_mode, mix_rate=48000.0)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([area], dtype=abi.area)
    fixed = np.zeros(1, dtype=abi.params)
    fixed["mix_volumes"][0, :, 0], fixed["mix_volumes"][0, :, 1] = 0.25, 0.5
    fixed["pitch_scale"], fixed["attenuation_filter_cutoff_hz"], fixed["update_parameters"], fixed["n_bus"] = 1.0, 5000.0, 1, 1
    fixed["bus_volumes"][0, 0, :, 0], fixed["bus_volumes"][0, 0, :, 1] = 0.25, 0.5
    inst = np.arange(V, dtype=np.int32)
    with orc.OracleMixer(**cfg) as o:
        o.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=mode_b))
        o.spatializer_set(1, abi.spatializer_defaults(mix_channel_mode=mode_b))
        o.instance_init(inst[:-1], 0)
        o.instance_init(inst[-1:], 1)
        voices = synth.make_voices(V)
        for b in range(blocks):
            if b == 0:
                o.voice_init(inst)
            o.gain_compute(emitters[b][:-1], listeners, areas, want_params=False)
            o.params_set(inst[-1:], fixed)
            if b == 0:
                o.instance_start(inst)
            want, _ = o.mix_block(voices, sources[b], F, want_peaks=False)
            assert np.array_equal(S.routing(got[b]), S.routing(want)), f"block {b}: routing differs"
            ok, worst, nbad = S.sample_close(got[b], want)
            assert ok, f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"


@pytest.mark.gpu
def test_per_call_virtuals_and_refusal_of_overriders(host_test):
    """process_frames / mix_channel with the reference's signature on the host mirror (audio_spatializer.h:146,148): built-in
    behaviour per call through the C ABI, wrong playback data is a no-op with an error, and an instance that overrides them
    (uses_builtin_dsp() == false) is refused by the batched step instead of being mixed wrongly."""
    out = subprocess.run([host_test, "percall"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "percall ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode_b", [0, 1])
def test_stream_lifecycle_through_batch_mixer_matches_oracle(host_test, orc, tmp_path, mode_b):
    """BatchMixer::mix_streams: the caller hands over what AudioStreamPlayback::mix returned; lookahead, end fade, tails and
    deactivation run on the device, finished playbacks leave the lists.  Against the oracle's stream form."""
    V, F, blocks, speaker_mode, num_buses = 24, 128, 10, abi.SPEAKER_SURROUND_51, 2
    rng = np.random.default_rng(9)
    length = rng.integers(F // 2, F * 5, size=V).astype(np.int32)
    length[0], length[1], length[2] = 3 * F, 40, F * (blocks + 3)
    emitters = [synth.make_emitters(V, block=b, dt=F / 48000.0) for b in range(blocks)]
    sources = [synth.make_sources(V, F, block=b, mix_rate=48000.0) for b in range(blocks)]
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    # the file carries, per block, the frames each stream delivers in that block at the START of the row
    pos = np.zeros(V, dtype=np.int64)
    rows = []
    for b in range(blocks):
        r = np.zeros((V, F, 2), dtype=np.float32)
        for v in range(V):
            n = int(min(F, max(0, length[v] - pos[v])))
            r[v, :n] = sources[b][v, :n]
            pos[v] += n
        rows.append(r)
    with open(inp, "wb") as f:
        f.write(np.array([V, F, blocks, speaker_mode, num_buses, mode_b], dtype=np.int32).tobytes())
        f.write(length.tobytes())
        for b in range(blocks):
            f.write(emitters[b].tobytes())
            f.write(rows[b].tobytes())
    out = subprocess.run([host_test, "stream", str(inp), str(outp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    C = speaker_mode + 1
    raw = np.fromfile(outp, dtype=np.uint8)
    per = num_buses * C * F * 2 * 4 + 4
    assert raw.size == blocks * per
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=num_buses, speaker_mode=speaker_mode, mix_rate=48000.0)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    inst = np.arange(V, dtype=np.int32)
    voices_all = synth.make_voices(V)
    with orc.OracleMixer(**cfg) as o:
        o.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=mode_b))
        o.instance_init(inst, 0)
        alive = np.ones(V, dtype=bool)
        pos[:] = 0
        for b in range(blocks):
            if b == 0:
                o.voice_init(inst)
            # the host layer computes gains for every instance, stopped or not
            em = emitters[b].copy()
            em["basis_z"] = 0
            em["basis_z"][:, 2] = 1.0  # host_test leaves the node basis at identity
            em["bus"], em["area"], em["pitch_scale"] = 0, -1, 1.0
            o.gain_compute(em, listeners, None, want_params=False)
            if b == 0:
                o.instance_start(inst)
            live = np.nonzero(alive)[0]
            mixed = np.array([int(min(F, max(0, length[v] - pos[v]))) for v in live], dtype=np.int32)
            for v, n in zip(live, mixed):
                pos[v] += n
            want, status = o.mix_block_stream(voices_all[live], rows[b], mixed, F)
            dead = live[(status & 1) == 0]
            alive[dead] = False
            if len(dead):
                o.instance_stop(dead.astype(np.int32))  # one playback per instance here: the instance stops with it
            blk = raw[b * per:(b + 1) * per]
            got = blk[:-4].view(np.float32).reshape(num_buses, C, F, 2)
            n_alive = int(blk[-4:].view(np.int32)[0])
            assert n_alive == int(alive.sum()), f"block {b}: {n_alive} playbacks alive, oracle says {int(alive.sum())}"
            assert np.array_equal(S.routing(got), S.routing(want)), f"block {b}: routing differs"
            ok, worst, nbad = S.sample_close(got, want)
            assert ok, f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"
        assert alive.sum() >= 1 and (~alive).sum() >= V // 2
