"""Voice lifecycle (reference audio_spatializer.cpp:353-408, :464-492): 64-frame lookahead splice, end-of-stream
fade (0.96^k x linear over the last 64 valid frames), zero-input tails, deactivation once a tail's block peak is at or below
playback_disable_threshold_db.

CPU: the oracle's stream form (orc_mix_block_stream) against the reference module's OWN _mix_from_playback_list
(oracle/_ref, fed through AudioStreamPlayback::mix returning short counts) — bit for bit, block by block, including the
block at which every voice is deactivated.
GPU: gas_mix_block_stream (lifecycle table on the device) against the oracle.
"""
import numpy as np
import pytest

import scenarios as S

abi, synth = S.abi, S.synth


def lifecycle_scenario(mode_b=True, speaker_mode=abi.SPEAKER_SURROUND_51, filt=True, voices=40, vpi=2, frames=256, blocks=14, seed=3):
    rng = np.random.default_rng(seed)
    # stream lengths in frames: some end inside a block, some exactly at a block edge, some shorter than the lookahead,
    # one empty stream, some outlive the scenario
    length = rng.integers(frames // 2, frames * (blocks - 6), size=voices)
    length[0] = 3 * frames            # ends exactly at a block edge
    length[1] = 3 * frames + 1
    length[2] = 40                    # shorter than the lookahead
    length[3] = 0                     # delivers nothing at all
    length[4] = 2 * frames + 63
    length[5] = 2 * frames + 64
    length[-1] = frames * (blocks + 5)  # still playing when the scenario ends
    start = np.zeros(voices, dtype=np.int64)
    start[7::9] = 2                   # late starters (SafeList order: newest first)
    return dict(mode_b=mode_b, speaker_mode=speaker_mode, filt=filt, voices=voices, vpi=vpi, frames=frames, blocks=blocks,
                length=length, start=start)


def run_stream(mixer, sc, threshold_db=-80.0, keep_alive=False):
    """keep_alive: one extra, silent, never-ending voice per instance (slots V .. V + instances - 1).  It adds exact zeros
    to every sum, so it changes no output bit; what it does is keep the reference instance's playback_active flag set in
    the block in which its last real voice is deactivated (see test_reference_drops_the_last_block_of_the_other_pairs)."""
    V, F, vpi = sc["voices"], sc["frames"], sc["vpi"]
    n_inst = (V + vpi - 1) // vpi
    inst = np.arange(n_inst, dtype=np.int32)
    spat = abi.spatializer_defaults(mix_channel_mode=int(sc["mode_b"]), **sc.get("spat", {}))
    mixer.spatializer_set(0, spat)
    mixer.instance_init(inst, 0)
    if threshold_db != -80.0:
        mixer.set_playback_disable_threshold_db(inst, threshold_db)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    voices_all = synth.make_voices(V, voices_per_instance=vpi)
    keep = synth.make_voices(n_inst, voice0=V, voices_per_instance=1)
    keep["src_row"] = V  # one shared row of zeros
    # the whole stream of every voice: continuous sources, cut at the voice's length
    stream = np.concatenate([synth.make_sources(V, F, block=b) for b in range(sc["blocks"])], axis=1)
    pos = np.zeros(V, dtype=np.int64)
    started = np.zeros(V, dtype=bool)
    alive = np.zeros(V, dtype=bool)
    out = dict(bus=[], status=[])
    for b in range(sc["blocks"]):
        em = synth.make_emitters(n_inst, block=b, dt=F / 48000.0)
        p = mixer.gain_compute(em, listeners, None)
        if not sc["filt"]:
            p = p.copy()
            p["linear_attenuation"] = 0.0
            mixer.params_set(inst, p)
        if b == 0:
            mixer.instance_start(inst)
            if keep_alive:
                mixer.voice_init(keep["voice"])
        fresh = (sc["start"] == b) & ~started
        if fresh.any():
            mixer.voice_init(np.nonzero(fresh)[0].astype(np.int32))
            started |= fresh
            alive |= fresh
        live = np.nonzero(alive)[0]
        src = np.zeros((V + (1 if keep_alive else 0), F, 2), dtype=np.float32)
        mixed = np.zeros(V, dtype=np.int32)
        for v in live:
            n = int(min(F, max(0, sc["length"][v] - pos[v])))
            src[v, :n] = stream[v, pos[v]:pos[v] + n]
            mixed[v] = n
            pos[v] += n
        voices = voices_all[live].copy()
        counts = mixed[live]
        if keep_alive:
            voices = np.concatenate([voices, keep])
            counts = np.concatenate([counts, np.full(n_inst, F, dtype=np.int32)])
        bus, status = mixer.mix_block_stream(voices, src, counts, F)
        full = np.zeros(V, dtype=np.int32)
        full[live] = status[: len(live)]
        alive &= (full & 1) == 1   # a deactivated voice is dropped from the list, like _manage_playback_state does
        out["bus"].append(bus)
        out["status"].append(full)
    out["state"] = mixer.voice_state_export(np.arange(V, dtype=np.int32))
    return out


def _bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("mode_b", [False, True])
@pytest.mark.parametrize("filt", [False, True])
def test_oracle_lifecycle_matches_reference_code(orc, mode_b, filt):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    sc = lifecycle_scenario(mode_b=mode_b, filt=filt)
    cfg = dict(max_instances=sc["voices"], max_voices=2 * sc["voices"], max_frames=sc["frames"], max_spatializers=2, num_buses=2,
               speaker_mode=sc["speaker_mode"], mix_rate=48000.0)
    with orc.OracleMixer(**cfg) as o, ref.RefMixer(**cfg) as r:
        e0 = r.error_count
        want = run_stream(o, sc)
        got = run_stream(r, sc, keep_alive=True)
        assert r.error_count == e0, ref.load().ref_last_error().decode()
    deactivated = 0
    for b in range(sc["blocks"]):
        assert np.array_equal(got["status"][b] & 1, want["status"][b] & 1), f"block {b}: active flags differ"
        assert _bits_equal(got["bus"][b], want["bus"][b]), \
            f"block {b}: bus differs, max {np.abs(got['bus'][b] - want['bus'][b]).max():.3e}"
        deactivated += int(((want["status"][b] & 1) == 0).sum())
    # the scenario exercises what it claims to: streams end, tails die, someone is still playing at the end
    assert (want["status"][-1] & 2).any(), "no voice still has frames at the end"
    assert ((want["status"][-1] & 1) == 0).sum() >= sc["voices"] // 2, "too few voices were deactivated"
    if filt:  # filtered tails ring below -80 dB for a few blocks before the voice is dropped
        ended = [int(np.argmax((np.stack(want["status"])[:, v] & 2) == 0)) for v in range(sc["voices"])]
        dropped = [int(np.argmax((np.stack(want["status"])[:, v] & 1) == 0)) for v in range(sc["voices"])]
        assert any(d > e for e, d in zip(ended, dropped) if d > 0), "no tail outlived its stream"


def test_reference_drops_the_last_block_of_the_other_pairs(orc):
    """A reference artefact the batched mixer deliberately does NOT reproduce.  When the last voice of a Mode-B instance is
    deactivated, _manage_playback_state stops the proxies and clears playback_active in the middle of the AudioServer mix
    step (audio_spatializer.cpp:484-491, reached from the first proxy's mix() call); the proxies AudioServer mixes after
    that one then deliver nothing (:683-690), so the final block (at or below -80 dBFS per voice) reaches only the pair
    whose proxy happened to be mixed first.  Which pair that is depends on AudioServer's playback order, which the module
    does not control.  The oracle and the CUDA path mix that block on every pair."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    sc = lifecycle_scenario(mode_b=True, filt=True, voices=8, vpi=1, blocks=8)
    sc["voices"], sc["blocks"] = 1, 3
    sc["length"], sc["start"] = np.array([300]), np.zeros(1, dtype=np.int64)
    cfg = dict(max_instances=1, max_voices=2, max_frames=sc["frames"], max_spatializers=2, num_buses=2, speaker_mode=sc["speaker_mode"],
               mix_rate=48000.0)
    with orc.OracleMixer(**cfg) as o, ref.RefMixer(**cfg) as r:
        want = run_stream(o, sc)
        got = run_stream(r, sc)
    last = int(np.argmax((np.stack(want["status"])[:, 0] & 1) == 0))  # block in which the voice is deactivated
    assert last > 0
    for b in range(last):
        assert _bits_equal(got["bus"][b], want["bus"][b])
    assert _bits_equal(got["bus"][last][:, 0], want["bus"][last][:, 0]), "the first proxy's pair still gets the block"
    assert np.abs(want["bus"][last][:, 1:]).max() > 0 and not got["bus"][last][:, 1:].any(), "the reference drops the other pairs"
    assert np.abs(want["bus"][last][:, 1:]).max() <= 2e-4, "what is dropped is a tail at or below the -80 dB threshold"


def test_oracle_lifecycle_threshold_property(orc):
    """A higher playback_disable_threshold_db drops tails sooner (never later)."""
    sc = lifecycle_scenario(mode_b=True, filt=True, voices=24, blocks=10)
    cfg = dict(max_instances=sc["voices"], max_voices=sc["voices"], max_frames=sc["frames"], max_spatializers=2, num_buses=2,
               speaker_mode=sc["speaker_mode"], mix_rate=48000.0)
    drops = []
    for db in (-80.0, -40.0):
        with orc.OracleMixer(**cfg) as o:
            out = run_stream(o, sc, threshold_db=db)
        st = np.stack(out["status"])
        drops.append(np.array([int(np.argmax((st[:, v] & 1) == 0)) if ((st[:, v] & 1) == 0).any() else 99 for v in range(sc["voices"])]))
    assert np.all(drops[1] <= drops[0]) and np.any(drops[1] < drops[0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode_b", [False, True])
@pytest.mark.parametrize("filt", [False, True])
def test_cuda_lifecycle_matches_oracle(gas, orc, mode_b, filt):
    sc = lifecycle_scenario(mode_b=mode_b, filt=filt)
    cfg = dict(max_instances=sc["voices"], max_voices=sc["voices"], max_frames=sc["frames"], max_spatializers=2, num_buses=2,
               speaker_mode=sc["speaker_mode"], mix_rate=48000.0)
    with orc.OracleMixer(**cfg) as o, gas.Mixer(**cfg) as m:
        want = run_stream(o, sc)
        got = run_stream(m, sc)
    for b in range(sc["blocks"]):
        assert np.array_equal(got["status"][b], want["status"][b]), f"block {b}: voice status differs"
        assert np.array_equal(S.routing(got["bus"][b]), S.routing(want["bus"][b])), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(got["bus"][b], want["bus"][b])
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"


def _nan_aware_bits_equal(a, b):
    """Bit for bit where both are numbers, NaN where either is (the payload of a NaN depends on operand order)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def nan_tail_scenario(mode_b):
    """SURVEY Q1 meets Q16: un-normalised directions into SPCAP with a non-integer tightness give NaN pan gains to some voices
    (audio_spatializer_3d.cpp:930); their streams end inside the run, so they go through the end fade and a zero-input tail block:
    0 x NaN is NaN, the bus stays poisoned for that block too, and the NaN samples never raise the peak (`l > peak.left` is false,
    audio_spatializer.cpp:436), so the voice is dropped after its first tail block."""
    sc = lifecycle_scenario(mode_b=mode_b, filt=False, voices=24, vpi=1, blocks=8)
    sc["spat"] = dict(panning_strength=1.5)
    return sc


@pytest.mark.parametrize("mode_b", [False, True])
def test_oracle_nan_gain_tails_match_reference_code(orc, mode_b):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    sc = nan_tail_scenario(mode_b)
    cfg = dict(max_instances=sc["voices"], max_voices=2 * sc["voices"], max_frames=sc["frames"], max_spatializers=2, num_buses=2,
               speaker_mode=sc["speaker_mode"], mix_rate=48000.0)
    with orc.OracleMixer(**cfg) as o, ref.RefMixer(**cfg) as r:
        want = run_stream(o, sc)
        got = run_stream(r, sc, keep_alive=True)
    poisoned_tail = False
    for b in range(sc["blocks"]):
        assert np.array_equal(got["status"][b] & 1, want["status"][b] & 1), f"block {b}: active flags differ"
        assert _nan_aware_bits_equal(got["bus"][b], want["bus"][b]), f"block {b}: bus differs"
        tails = (want["status"][b - 1] & 3) == 1 if b else np.zeros(sc["voices"], dtype=bool)  # active without frames when the block began
        poisoned_tail = poisoned_tail or (bool(tails.any()) and bool(np.isnan(want["bus"][b]).any()))
    assert poisoned_tail, "the scenario never mixed a zero-input tail while the bus was NaN"


@pytest.mark.gpu
@pytest.mark.parametrize("mode_b", [False, True])
def test_cuda_nan_gain_tails_match_oracle(gas, orc, mode_b):
    sc = nan_tail_scenario(mode_b)
    cfg = dict(max_instances=sc["voices"], max_voices=sc["voices"], max_frames=sc["frames"], max_spatializers=2, num_buses=2,
               speaker_mode=sc["speaker_mode"], mix_rate=48000.0)
    with orc.OracleMixer(**cfg) as o, gas.Mixer(**cfg) as m:
        want = run_stream(o, sc)
        got = run_stream(m, sc)
    for b in range(sc["blocks"]):
        assert np.array_equal(got["status"][b], want["status"][b]), f"block {b}: voice status differs"
        assert np.array_equal(np.isnan(got["bus"][b]), np.isnan(want["bus"][b])), f"block {b}: NaN pattern differs"
        ok, worst, nbad = S.sample_close(got["bus"][b], want["bus"][b])
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"
