"""Pins the restated CPU oracle (oracle/gas_oracle.c) against the reference module's OWN code.

oracle/_ref/libgas_ref.so is /root/reference/*.cpp, unmodified, compiled against the godot-lite stand-in
headers (oracle/godot_lite/) and driven by oracle/ref_harness.cpp through the same call sequence as the
oracle.  Every comparison here is BIT-EXACT (float32 bit patterns, NaNs included) unless the reference leaves
the float summation order to the start order of voices (noted where it applies).

What this pins: every module-side line of the path (audio_spatializer_3d.cpp:57-609,903-938,
audio_spatializer.cpp:258-471, spatializer_parameters.cpp, audio_spatializer_effect.cpp:33-88,
audio_stream_player_spatial.cpp:405-413).  What it cannot pin: upstream Godot itself (AudioFilterSW,
AudioServer::_mix_step, Math::*, Basis/Transform3D) — both sides restate it from Godot 4.x as recalled, in two
independently written restatements (C in the oracle, C++ in godot_lite) that are checked against each other.
"""
import ctypes as C

import numpy as np
import pytest

import scenarios as S
from oracle import ref as _ref

pytestmark = pytest.mark.skipif(not _ref.available(), reason="oracle/_ref not built and /root/reference not present")
abi = S.abi
MODES = {"A": 0, "B": 1}


@pytest.fixture(scope="module")
def ref():
    _ref.load()
    return _ref


def _bits(a):
    return np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint8)


def _same_bits(got, want, what):
    """float32 arrays equal bit for bit; NaNs must sit in the same places (payload / sign bits of a NaN are
    not compared: they depend on which operand the SSE instruction propagated)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    if got.dtype.kind == "f":
        ng, nw = np.isnan(got), np.isnan(want)
        assert np.array_equal(ng, nw), f"{what}: NaN pattern differs"
        g = np.where(ng, 0, got).astype(got.dtype)
        w = np.where(nw, 0, want).astype(want.dtype)
        same = g.view(np.uint32) == w.view(np.uint32) if got.dtype == np.float32 else g == w
        if not same.all():
            d = np.abs(g.astype(np.float64) - w.astype(np.float64)).max()
            raise AssertionError(f"{what}: {int((~same).sum())} of {same.size} values differ, max abs diff {d:.3e}")
    else:
        assert np.array_equal(got, want), what


def _play(orc, ref, sc):
    cfg = S.config_of(sc)
    with orc.OracleMixer(**cfg) as o, ref.RefMixer(**cfg) as r:
        e0 = r.error_count
        want = S.run(o, sc)
        got = S.run(r, sc)
        assert r.error_count == e0, f"{sc['name']}: the reference logged an error: {ref.load().ref_last_error().decode()}"
    return got, want


def _check(got, want, sc, exact_bus=True, state=True):
    assert len(got["params"]) == len(want["params"]) > 0
    for b, (pg, pw) in enumerate(zip(got["params"], want["params"])):
        for f in pg.dtype.names:
            _same_bits(pg[f], pw[f], f"{sc['name']} block {b} params.{f}")
    for b, (bg, bw) in enumerate(zip(got["bus"], want["bus"])):
        assert np.isnan(bw).any() or np.abs(bw).max() > 0, "scenario mixes nothing"
        if exact_bus:
            _same_bits(bg, bw, f"{sc['name']} block {b} bus")
        else:
            assert np.array_equal(S.routing(bg), S.routing(bw))
            ok, worst, nbad = S.sample_close(bg, bw, rel=1e-6, abs_tol=1e-8)
            assert ok, f"{sc['name']} block {b}: {nbad} samples differ beyond summation order, worst {worst:.3e}"
    if state:
        for f in ("prev_mix_volumes", "filter_processors", "effect_history"):
            a, b_ = got["state"][f], want["state"][f]
            assert np.array_equal(_bits(a), _bits(b_)), f"{sc['name']}: final voice state {f} differs"


# ---- whole-path scenarios: gains -> bus map -> per-voice mix -> AudioServer accumulate ------------------------------


@pytest.mark.parametrize("speakers", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", ["A", "B"])
@pytest.mark.parametrize("filt", [False, True])
def test_mix_path_all_speaker_modes(orc, ref, mode, speakers, filt):
    """process_frames / mix_channel + get_bus_map + the bus accumulate, filter off (forced) and on (Q11-Q15)."""
    sc = S.default_scenario(name=f"{mode}-{speakers}-{filt}", voices=48, speaker_mode=speakers, spat=dict(mix_channel_mode=MODES[mode]),
                            force_filter_off=not filt, blocks=3)
    _check(*_play(orc, ref, sc), sc)


@pytest.mark.parametrize("mode", ["A", "B"])
@pytest.mark.parametrize("uniformity", [0.0, 0.6])
def test_reverb_area_override_bus_two_listeners(orc, ref, mode, uniformity):
    """calc_reverb_vol (Q17), area override bus + reverb bus (Q18), max-combine over two listeners (Q7)."""
    sc = S.default_scenario(name=f"reverb-{mode}-{uniformity}", voices=40, speaker_mode=abi.SPEAKER_SURROUND_71, num_buses=3,
                            spat=dict(mix_channel_mode=MODES[mode]), listeners="two", area_fraction=0.5,
                            area=dict(reverb_bus=1, amount=0.5, uniformity=uniformity, override_bus=True, bus=2), blocks=3)
    _check(*_play(orc, ref, sc), sc)
    sc = S.default_scenario(name=f"reverb-stereo-{mode}-{uniformity}", voices=24, speaker_mode=abi.SPEAKER_MODE_STEREO,
                            spat=dict(mix_channel_mode=MODES[mode]), area_fraction=0.5, force_filter_off=True,
                            area=dict(reverb_bus=1, amount=0.7, uniformity=uniformity), blocks=2)
    _check(*_play(orc, ref, sc), sc)


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_attenuation_models_max_distance_emission_angle(orc, ref, model):
    """get_attenuation_db (Q2-Q4), max_distance taper and skip (Q5), emission-angle filter gain, last listener wins (Q6)."""
    sc = S.default_scenario(name=f"model-{model}", voices=33, speaker_mode=abi.SPEAKER_SURROUND_71, listeners="two", blocks=2,
                            spat=dict(attenuation_model=model, max_distance=100.0, mix_channel_mode=1, emission_angle_enabled=1, emission_angle=30.0))
    _check(*_play(orc, ref, sc), sc)


@pytest.mark.parametrize("tracking", [1, 2])
def test_doppler_pitch(orc, ref, tracking):
    """audio_spatializer_3d.cpp:405-434 with moving emitters and moving, rotated listeners."""
    listeners = [S.synth.rotated_listener(velocity=(1.0, -2.0, 0.5)), S.synth.rotated_listener(yaw=-1.1, origin=(-5, 0, 2), velocity=(0, 0, 0))]
    sc = S.default_scenario(name=f"doppler-{tracking}", voices=50, speaker_mode=abi.SPEAKER_SURROUND_51, listeners=listeners, blocks=2,
                            spat=dict(mix_channel_mode=1, doppler_tracking=tracking, doppler_speed_of_sound=200.0))
    got, want = _play(orc, ref, sc)
    _check(got, want, sc)
    pitch = want["params"][0]["pitch_scale"]
    assert np.ptp(pitch) > 1e-3, "Doppler did not move the pitch"


@pytest.mark.parametrize("speakers", [1, 2, 3])
@pytest.mark.parametrize("strength", [0.37, 1.3, 2.5])
def test_q1_non_integer_tightness_gives_nan_like_the_reference(orc, ref, speakers, strength):
    """Q1: the module hands the UN-normalised local position to SPCAP, so 1 + dir.src goes negative for sources behind
    a speaker and pow(negative, non-integer tightness) is NaN (audio_spatializer_3d.cpp:391 -> :930).  The oracle must
    produce NaN exactly where the reference does."""
    sc = S.default_scenario(name=f"q1-{strength}-{speakers}", voices=64, speaker_mode=speakers, blocks=2,
                            spat=dict(mix_channel_mode=1, panning_strength=strength))
    got, want = _play(orc, ref, sc)
    _check(got, want, sc)
    assert np.isnan(want["params"][0]["mix_volumes"]).any(), "scenario did not reach the NaN case"
    # integer tightness with the same geometry: negative base, finite (wrong-sign gain squared)
    sc = S.default_scenario(name=f"q1-int-{speakers}", voices=64, speaker_mode=speakers, blocks=2, spat=dict(mix_channel_mode=1, panning_strength=3.0))
    got, want = _play(orc, ref, sc)
    _check(got, want, sc)
    assert np.isfinite(want["params"][0]["mix_volumes"]).all()


@pytest.mark.parametrize("stages", [1, 2, 4])
def test_effect_chain(orc, ref, stages):
    """AudioSpatializerInstanceEffect::process_frames ping-pong chain of AudioEffectFilter instances, gain bound to the
    computed high-shelf gain by the (scripted) _process_effects hook, Mode A sends (Q15) to three buses."""
    chain = [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=stages)]
    sc = S.default_scenario(name=f"effect-{stages}", voices=37, speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=3, effect_chain=chain,
                            effect_gain_binding=0, area=dict(reverb_bus=2, amount=0.4), area_fraction=0.5, blocks=3)
    _check(*_play(orc, ref, sc), sc)


def test_effect_chain_three_effects_mixed_modes(orc, ref):
    chain = [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=2),
             dict(mode=abi.FILTER_LOWPASS, cutoff_hz=9000.0, resonance=0.7, gain=1.0, stages=1),
             dict(mode=abi.FILTER_PEAK, cutoff_hz=900.0, resonance=0.7, gain=2.0, stages=3)]
    sc = S.default_scenario(name="effect-3fx", voices=21, speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=2, effect_chain=chain, blocks=3)
    _check(*_play(orc, ref, sc), sc)


def test_polyphony_and_ragged_sizes(orc, ref):
    """Several voices per instance (the instance pre-sum), odd voice counts, non-power-of-two blocks."""
    sc = S.default_scenario(name="poly", voices=45, voices_per_instance=3, frames=480, speaker_mode=abi.SPEAKER_SURROUND_51,
                            spat=dict(mix_channel_mode=1), blocks=3)
    _check(*_play(orc, ref, sc), sc)
    sc = S.default_scenario(name="poly-A", voices=44, voices_per_instance=4, frames=130, speaker_mode=abi.SPEAKER_MODE_STEREO,
                            spat=dict(mix_channel_mode=0), blocks=3)
    _check(*_play(orc, ref, sc), sc)


def test_late_start_and_silent_voices(orc, ref):
    """Voices joining at block 2 and zero-input tail voices.  A voice that starts later is inserted at the HEAD of the
    module's SafeList (audio_spatializer.cpp:74), so the per-instance float sum meets the newest voice first; the oracle
    keeps the start order of its voices and sums in the same order, so this too is bit for bit."""
    sc = S.default_scenario(name="late", voices=45, voices_per_instance=3, frames=480, speaker_mode=abi.SPEAKER_SURROUND_51,
                            spat=dict(mix_channel_mode=1), blocks=4, start_late=2, silent_every=7)
    _check(*_play(orc, ref, sc), sc)
    sc = S.default_scenario(name="late-A", voices=45, voices_per_instance=5, frames=130, speaker_mode=abi.SPEAKER_MODE_STEREO,
                            spat=dict(mix_channel_mode=0), blocks=4, start_late=1, silent_every=4)
    _check(*_play(orc, ref, sc), sc)


# ---- scalar pieces ---------------------------------------------------------------------------------------------------


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_get_attenuation_db_bitwise(orc, ref):
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(7)
    for model in range(4):
        s = abi.spatializer_defaults(attenuation_model=model, unit_size=float(rng.uniform(0.5, 30)))
        sp = np.ascontiguousarray(s).reshape(1)
        for d in list(rng.uniform(0, 300, 40)) + [0.0, 1e-7, 1e6]:
            vdb, mdb = float(rng.uniform(-30, 6)), float(rng.uniform(-6, 6))
            a = np.float32(lo.orc_get_attenuation_db(_fp(sp), vdb, mdb, float(d)))
            b = np.float32(lr.ref_get_attenuation_db(_fp(sp), vdb, mdb, float(d)))
            assert a.view(np.uint32) == b.view(np.uint32), (model, d, a, b)


def test_calc_output_vol_bitwise(orc, ref):
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(11)
    for mode in range(4):
        for k in range(60):
            d = (rng.normal(size=3) * rng.choice([0.01, 1.0, 50.0])).astype(np.float32)
            if k == 0:
                d[:] = 0
            if k == 1:
                d[:] = (0, 5, 0)
            ps = float(rng.choice([0.0, 0.5, 1.0, 2.0, 3.0, 0.77]))
            a = np.full((4, 2), 7.0, np.float32)
            b = np.full((4, 2), 7.0, np.float32)
            lo.orc_calc_output_vol(mode, 0.5, ps, _fp(d), _fp(a))
            lr.ref_calc_output_vol(mode, 0.5, ps, _fp(d), _fp(b))
            _same_bits(b, a, f"calc_output_vol mode {mode} dir {d} ps {ps}")


def test_spcap_bitwise(orc, ref):
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(13)
    for count in (3, 5, 7):
        eff_o = np.zeros(7, np.float32)
        lo.orc_spcap_effective_speakers(count, _fp(eff_o))
        for _ in range(40):
            d = rng.normal(size=3).astype(np.float32)
            d /= np.float32(np.linalg.norm(d))
            t = float(rng.choice([1.0, 2.0, 0.5, 1.7, 4.0]))
            vo, vr, er = np.zeros(7, np.float32), np.zeros(7, np.float32), np.zeros(7, np.float32)
            lo.orc_spcap_calculate(count, _fp(d), t, _fp(vo))
            lr.ref_spcap_calculate(count, _fp(d), t, _fp(vr), _fp(er))
            _same_bits(vr, vo, f"spcap volumes count {count}")
            _same_bits(er[:count], eff_o[:count], f"spcap effective speakers count {count}")


def test_get_bus_map_bitwise(orc, ref):
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(17)
    for trial in range(30):
        p = np.zeros(1, dtype=abi.params)
        p["mix_volumes"] = rng.uniform(0, 1, (4, 2)).astype(np.float32)
        if trial % 3 == 0:
            p["mix_volumes"][0, 1, 0] = 0.0  # mix_vol <= 0 -> send 0 (audio_spatializer.cpp:304-309)
        n = int(rng.integers(1, 7))
        p["n_bus"] = n
        p["bus"][0, :n] = rng.permutation(8)[:n]
        p["bus_volumes"] = rng.uniform(0, 1, (6, 4, 2)).astype(np.float32)
        for mc in (0, 1):
            for ch in range(4):
                bo, br = np.zeros(6, np.int32), np.zeros(6, np.int32)
                vo, vr = np.zeros((6, 4, 2), np.float32), np.zeros((6, 4, 2), np.float32)
                no = lo.orc_get_bus_map(_fp(p), mc, ch, _fp(bo), _fp(vo))
                nr = lr.ref_get_bus_map(_fp(p), mc, ch, _fp(br), _fp(vr))
                assert no == nr == n
                assert np.array_equal(bo[:n], br[:n])
                _same_bits(vr[:n], vo[:n], "bus map volumes")


def test_filter_restatements_agree(orc, ref):
    """Upstream AudioFilterSW::prepare_coefficients exists twice (oracle C, godot-lite C++), written independently from
    the recalled upstream source: they must agree bit for bit, all eight modes."""
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(19)
    for mode in range(8):
        for _ in range(25):
            cutoff, res = float(rng.uniform(20, 30000)), float(rng.choice([0.0, 0.5, 1.0, 2.3]))
            gain, stages = float(rng.choice([0.0005, 0.05, 0.3, 1.0, 2.0])), int(rng.integers(1, 5))
            sr = float(rng.choice([44100.0, 48000.0]))
            a, b = np.zeros(5, np.float32), np.zeros(5, np.float32)
            lo.orc_filter_prepare_coefficients(mode, cutoff, res, gain, stages, sr, _fp(a))
            lr.ref_filter_prepare_coefficients(mode, cutoff, res, gain, stages, sr, _fp(b))
            _same_bits(b, a, f"prepare_coefficients mode {mode}")


@pytest.mark.parametrize("gain", [0.0, 0.0009, 0.001, 0.25, 1.0])
def test_process_frames_and_mix_channel_single_voice(orc, ref, gain):
    """One voice through process_frames / mix_channel over three blocks with carried state: the filter threshold
    (>= 0.001, Q11), fade-in of the coefficients from zero on the first block (Q12), prev-volume bookkeeping (Q14)."""
    lo, lr = orc.load(), ref.load()
    rng = np.random.default_rng(23)
    F = 256
    so, sr_ = np.zeros(1, dtype=abi.voice_state), np.zeros(1, dtype=abi.voice_state)
    for blk in range(3):
        p = np.zeros(1, dtype=abi.params)
        p["mix_volumes"] = rng.uniform(0, 1, (4, 2)).astype(np.float32)
        p["linear_attenuation"] = gain
        p["attenuation_filter_cutoff_hz"] = 5000.0
        src = rng.uniform(-0.5, 0.5, (F, 2)).astype(np.float32)
        for ch in range(4):
            oo, orf = np.zeros((F, 2), np.float32), np.zeros((F, 2), np.float32)
            lo.orc_mix_channel_3d(_fp(p), _fp(so), 48000.0, ch, _fp(oo), _fp(src), F)
            lr.ref_mix_channel_3d(_fp(p), _fp(sr_), 48000.0, ch, _fp(orf), _fp(src), F)
            _same_bits(orf, oo, f"mix_channel block {blk} ch {ch}")
        assert np.array_equal(_bits(so), _bits(sr_))
    so, sr_ = np.zeros(1, dtype=abi.voice_state), np.zeros(1, dtype=abi.voice_state)
    for blk in range(3):
        p = np.zeros(1, dtype=abi.params)
        p["mix_volumes"] = rng.uniform(0, 1, (4, 2)).astype(np.float32)
        p["linear_attenuation"] = gain
        p["attenuation_filter_cutoff_hz"] = 3000.0
        src = rng.uniform(-0.5, 0.5, (F, 2)).astype(np.float32)
        oo, orf = np.zeros((F, 2), np.float32), np.zeros((F, 2), np.float32)
        lo.orc_process_frames_3d(_fp(p), _fp(so), 44100.0, _fp(oo), _fp(src), F)
        lr.ref_process_frames_3d(_fp(p), _fp(sr_), 44100.0, _fp(orf), _fp(src), F)
        _same_bits(orf, oo, f"process_frames block {blk}")
        assert np.array_equal(_bits(so), _bits(sr_))


def test_reference_validation_and_registration(ref):
    """The module's own setters reject what gas_spatializer_set rejects (audio_spatializer_3d.cpp:670-760), and its
    registration entry point ran (register_types.cpp:40-60: 12 classes)."""
    lib = ref.load()
    with ref.RefMixer(max_voices=4, max_instances=4) as r:
        assert lib.ref_registered_classes() == 12
        for bad in (dict(max_distance=-1.0), dict(emission_angle=91.0), dict(panning_strength=-0.1), dict(doppler_speed_of_sound=0.0)):
            with pytest.raises(ref.RefError):
                r.spatializer_set(0, abi.spatializer_defaults(**bad))
        r.spatializer_set(0, abi.spatializer_defaults())
