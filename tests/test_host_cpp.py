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
