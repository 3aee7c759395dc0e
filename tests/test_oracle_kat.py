"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors for this path, so these answers are
SELF-DERIVED by hand from the cited reference lines (file:line into the reference tree); they are the
pins the oracle is trusted on.  No GPU needed.
"""
import ctypes as C
import math

import numpy as np
import pytest

import scenarios as S

abi = S.abi


def f32(x):
    return np.float32(x)


@pytest.fixture(scope="module")
def lib(orc):
    return orc.load()


def _spat(**kw):
    return abi.spatializer_defaults(**kw).reshape(1)


def _p(a):
    return C.c_void_p(a.ctypes.data)


# ---- Math::db_to_linear / linear_to_db -----------------------------------------------
def test_db_linear_roundtrip(lib):
    assert lib.orc_db_to_linear_f(0.0) == 1.0
    assert abs(lib.orc_db_to_linear_f(-20.0) - 0.1) < 1e-7
    assert abs(lib.orc_db_to_linear_f(-80.0) - 1e-4) < 1e-10  # playback_disable_threshold_db, audio_spatializer.h:87
    assert lib.orc_linear_to_db_f(1.0) == 0.0
    assert abs(lib.orc_linear_to_db_f(0.5) - (-6.0206)) < 1e-4


# ---- get_attenuation_db, audio_spatializer_3d.cpp:123-151 ------------------------------------------------
def test_attenuation_models(lib):
    unit = 10.0
    s = _spat(attenuation_model=abi.ATTENUATION_INVERSE_DISTANCE, unit_size=unit)
    # :127  linear_to_db(1 / (d/unit + 1e-5)) at d == unit  ->  -20*log10(1.00001)
    assert abs(lib.orc_get_attenuation_db(_p(s), 0.0, 3.0, unit) - (-20 * math.log10(1.00001))) < 1e-6
    assert abs(lib.orc_get_attenuation_db(_p(s), 0.0, 3.0, 2 * unit) - (-20 * math.log10(2.00001))) < 1e-5
    s = _spat(attenuation_model=abi.ATTENUATION_INVERSE_SQUARE_DISTANCE, unit_size=unit)
    assert abs(lib.orc_get_attenuation_db(_p(s), 0.0, 3.0, 2 * unit) - (-20 * math.log10(4.00001))) < 1e-5  # :130-132
    s = _spat(attenuation_model=abi.ATTENUATION_LOGARITHMIC, unit_size=unit)
    # :135 natural log (Q2): d = e*unit -> -20
    assert abs(lib.orc_get_attenuation_db(_p(s), 0.0, 3.0, math.e * unit) - (-20.0)) < 1e-4
    s = _spat(attenuation_model=abi.ATTENUATION_DISABLED)
    assert lib.orc_get_attenuation_db(_p(s), -7.5, 3.0, 123.0) == -7.5  # :137 + :145
    # :145-148 volume_db added, then the SUM is clamped to max_db (Q4)
    s = _spat(attenuation_model=abi.ATTENUATION_INVERSE_DISTANCE, unit_size=unit)
    assert lib.orc_get_attenuation_db(_p(s), 10.0, 3.0, 0.01) == 3.0
    assert abs(lib.orc_get_attenuation_db(_p(s), -6.0, 3.0, unit) - (-6.0 - 20 * math.log10(1.00001))) < 1e-5


# ---- calc_output_vol_stereo, audio_spatializer_3d.cpp:103-110 (Q8) ---------------------------------------
def test_stereo_pan(lib):
    out = np.zeros((4, 2), np.float32)

    def pan(d, ps):
        out[:] = 0
        dd = np.array(d, np.float32)
        lib.orc_calc_output_vol_stereo(_p(dd), ps, _p(out))
        return out[0].copy()

    assert np.allclose(pan((1, 0, 0), 1.0), (0.0, 1.0), atol=1e-7)      # g=0, f=1, cosx=1
    assert np.allclose(pan((-1, 0, 0), 1.0), (1.0, 0.0), atol=1e-7)
    assert np.allclose(pan((0, 0, -1), 1.0), (math.sqrt(0.5),) * 2, atol=1e-7)
    assert np.allclose(pan((0, 5, 0), 1.0), (math.sqrt(0.5),) * 2, atol=1e-7)  # flatrad == 0 -> 1 (:107)
    # pan_strength 0.5: g=.25, f=.6 -> L = sqrt((1-.6)/2), R = sqrt((1+.6)/2)
    assert np.allclose(pan((3, 0, 0), 0.5), (math.sqrt(0.2), math.sqrt(0.8)), atol=1e-7)
    # scale invariance (Q1 does not affect stereo)
    assert np.array_equal(pan((0.3, 0.1, -0.4), 0.5), pan((0.6, 0.2, -0.8), 0.5))
    # only pair 0 is written
    assert not out[1:].any()
    # power complementarity L^2 + R^2 == 1
    l, r = pan((0.3, 0.0, -0.9), 0.7)
    assert abs(l * l + r * r - 1.0) < 1e-6


# ---- SPCAP, audio_spatializer_3d.cpp:47-55, :903-938 (Q9, Q10) --------------------------------------------
def test_spcap(lib):
    eff = np.zeros(7, np.float32)
    lib.orc_spcap_effective_speakers(3, _p(eff))
    r = math.sqrt(0.5)
    # FL: 0.5(1+1) + 0.5(1+FL.FR=0) + 0.5(1+FL.C=r)
    assert abs(eff[0] - (1.0 + 0.5 + 0.5 * (1 + r))) < 1e-6
    assert abs(eff[2] - (1.0 + (1 + r))) < 1e-6  # C: itself + 2 * 0.5(1+r)
    vol = np.zeros(7, np.float32)
    d = np.array((0, 0, -1), np.float32)
    for n in (3, 5, 7):
        vol[:] = 0
        lib.orc_spcap_calculate(n, _p(d), 1.0, _p(vol))
        assert abs(float((vol[:n].astype(np.float64) ** 2).sum()) - 1.0) < 1e-6  # sqrt(sq/sum) normalisation
        assert abs(vol[0] - vol[1]) < 1e-7                                        # symmetric source
        if n >= 5:
            assert abs(vol[3] - vol[4]) < 1e-7
        if n == 7:
            assert abs(vol[5] - vol[6]) < 1e-7
        assert vol[2] > vol[0]                                                    # centre speaker dominates
    out = np.zeros((4, 2), np.float32)
    lib.orc_calc_output_vol_surround(abi.SPEAKER_SURROUND_51, _p(d), 1.0, _p(out))
    assert out[1, 1] == 1.0 and not out[3].any()   # LFE always 1.0 (:91); side pair untouched in 5.1
    # Q1: the direction is NOT normalised, a longer vector changes the surround gains
    out2 = np.zeros((4, 2), np.float32)
    d2 = np.array((0.4, 0, -1.8), np.float32)
    d2n = d2 / np.linalg.norm(d2)
    lib.orc_calc_output_vol_surround(abi.SPEAKER_SURROUND_71, _p(d2), 1.0, _p(out))
    lib.orc_calc_output_vol_surround(abi.SPEAKER_SURROUND_71, _p(d2n.astype(np.float32)), 1.0, _p(out2))
    assert not np.allclose(out, out2, atol=1e-3)


# ---- AudioFilterSW::prepare_coefficients ------------------------------------------------
def _coeffs(lib, mode, cutoff, res, gain, stages=1, sr=48000.0):
    c = np.zeros(5, np.float32)
    lib.orc_filter_prepare_coefficients(mode, cutoff, res, gain, stages, sr, _p(c))
    return c.astype(np.float64)


def _gain_at(c, z):
    b0, b1, b2, a1, a2 = c  # feedback stored negated: y = b0 x + b1 x1 + b2 x2 + a1 y1 + a2 y2
    return (b0 + b1 / z + b2 / z ** 2) / (1 - a1 / z - a2 / z ** 2)


def test_filter_coefficients(lib):
    for gain in (0.0631, 0.25, 1.0):
        c = _coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, gain)
        assert abs(_gain_at(c, 1.0) - 1.0) < 1e-5            # high shelf: unity at DC
        assert abs(_gain_at(c, -1.0) - max(gain, 0.001) ** 2) < 1e-5  # A = gain (not sqrt): gain^2 at Nyquist
    c = _coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, 1.0)
    assert abs(c[0] - 1.0) < 1e-7 and abs(c[1] + c[3]) < 1e-7 and abs(c[2] + c[4]) < 1e-7  # gain 1 => identity
    c = _coeffs(lib, abi.FILTER_LOWPASS, 2000.0, 0.7, 1.0)
    assert abs(_gain_at(c, 1.0) - 1.0) < 1e-5 and abs(_gain_at(c, -1.0)) < 1e-6
    c = _coeffs(lib, abi.FILTER_HIGHPASS, 2000.0, 0.7, 1.0)
    assert abs(_gain_at(c, -1.0) - 1.0) < 1e-5 and abs(_gain_at(c, 1.0)) < 1e-6
    c = _coeffs(lib, abi.FILTER_LOWSHELF, 500.0, 1.0, 0.5)
    assert abs(_gain_at(c, 1.0) - 0.25) < 1e-5 and abs(_gain_at(c, -1.0) - 1.0) < 1e-5
    # gain floor 0.001 and cutoff clamp to sr/2 + 512
    assert np.allclose(_coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, 0.0), _coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, 0.001), rtol=1e-6)
    assert np.array_equal(_coeffs(lib, abi.FILTER_HIGHSHELF, 1e6, 1.0, 0.5), _coeffs(lib, abi.FILTER_HIGHSHELF, 24512.0, 1.0, 0.5))
    # stages > 1: tmpgain = gain^(1/(stages+1)) = 0.5 -> each stage has 0.25 at Nyquist
    c = _coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, 0.125, stages=2)
    assert abs(_gain_at(c, -1.0) - 0.25) < 1e-5


# ---- mix_channel / process_frames, audio_spatializer_3d.cpp:491-609 ------------------------------------------
def _params(mix, att=0.0, cutoff=5000.0):
    p = np.zeros(1, dtype=abi.params)
    p["mix_volumes"][0] = mix
    p["linear_attenuation"] = att
    p["attenuation_filter_cutoff_hz"] = cutoff
    p["pitch_scale"] = 1.0
    return p


def test_mix_channel_ramp_endpoints(lib):
    """Q13: t = i/F never reaches 1; out[0] == prev*src[0]; prev <- new afterwards (:589-608)."""
    F = 512
    src = np.ones((F, 2), np.float32)
    out = np.zeros((F, 2), np.float32)
    st = np.zeros(1, dtype=abi.voice_state)
    st["prev_mix_volumes"][0, 0] = (0.25, 0.5)
    p = _params([[1.0, 0.0], [0, 0], [0, 0], [0, 0]])
    lib.orc_mix_channel_3d(_p(p), _p(st), 48000.0, 0, _p(out), _p(src), F)
    assert out[0, 0] == f32(0.25) and out[0, 1] == f32(0.5)
    t = f32(511) / f32(512)
    assert out[-1, 0] == f32(f32(1.0) * t + (f32(1) - t) * f32(0.25))
    assert out[-1, 0] < 1.0
    assert np.all(np.diff(out[:, 0]) > 0) and np.all(np.diff(out[:, 1]) < 0)
    assert tuple(st["prev_mix_volumes"][0, 0]) == (1.0, 0.0)
    assert not st["filter_processors"]["b0"].any()  # filter untouched while linear_attenuation < 0.001 (:568)


def test_filter_fade_in_from_zero_and_identity(lib):
    """Q12: a fresh Processor has zero coefficients, so the first filtered block fades in from silence;
    Q11: with gain 1 the converged high shelf is the identity."""
    F = 512
    rng = np.random.default_rng(1)
    src = rng.uniform(-0.5, 0.5, (F, 2)).astype(np.float32)
    out = np.zeros((F, 2), np.float32)
    st = np.zeros(1, dtype=abi.voice_state)
    p = _params([[0.7, 0.7], [0, 0], [0, 0], [0, 0]], att=1.0)
    lib.orc_process_frames_3d(_p(p), _p(st), 48000.0, _p(out), _p(src), F)
    assert out[0, 0] == 0.0 and out[0, 1] == 0.0          # b0 == 0 on the first sample
    assert abs(out[-1, 0] - src[-1, 0]) < 1e-2            # nearly converged at the end of the block
    target = _coeffs(lib, abi.FILTER_HIGHSHELF, 5000.0, 1.0, 1.0)
    got = np.array([st["filter_processors"][k][0, 0] for k in ("b0", "b1", "b2", "a1", "a2")], np.float64)
    assert np.allclose(got, target, atol=2e-5)             # 512 float increments land next to the target
    assert tuple(st["prev_mix_volumes"][0, 0]) == (f32(0.7), f32(0.7))  # Q14: pair holding the max component
    for _ in range(3):
        lib.orc_process_frames_3d(_p(p), _p(st), 48000.0, _p(out), _p(src), F)
    assert np.allclose(out[8:], src[8:], atol=2e-4)        # identity once converged


def test_mode_a_prev_volume_pair_selection(lib):
    """Q14 (:537-551): strictly-greater scan, first pair wins ties, all-zero volumes keep index 0."""
    F = 8
    src = np.zeros((F, 2), np.float32)
    out = np.zeros((F, 2), np.float32)
    st = np.zeros(1, dtype=abi.voice_state)
    p = _params([[0.1, 0.2], [0.05, 0.9], [0.9, 0.1], [0.3, 0.3]])
    lib.orc_process_frames_3d(_p(p), _p(st), 48000.0, _p(out), _p(src), F)
    assert tuple(st["prev_mix_volumes"][0, 0]) == (f32(0.05), f32(0.9))


# ---- get_bus_map, audio_spatializer.cpp:274-324 (Q15) ----------------------------------------------------------
def test_get_bus_map(lib):
    p = np.zeros(1, dtype=abi.params)
    p["mix_volumes"][0] = [[0.5, 0.25], [0.0, 0.4], [0.2, 0.2], [0.1, 0.3]]
    p["n_bus"] = 2
    p["bus"][0, :2] = (0, 3)
    p["bus_volumes"][0, 0] = p["mix_volumes"][0]
    p["bus_volumes"][0, 1] = p["mix_volumes"][0] * np.float32(0.5)
    bus = np.zeros(6, np.int32)
    vol = np.zeros((6, 4, 2), np.float32)
    n = lib.orc_get_bus_map(_p(p), 1, 1, _p(bus), _p(vol))  # Mode B, proxy of pair 1
    assert n == 2 and tuple(bus[:2]) == (0, 3)
    assert vol[0, 1, 1] == 1.0 and vol[1, 1, 1] == 0.5      # main bus exactly 1, reverb bus the ratio
    assert vol[0, 1, 0] == 0.0                               # mix volume <= 0 => 0 (:304)
    assert not vol[:, [0, 2, 3]].any()                       # masked to the requested pair (:300)
    n = lib.orc_get_bus_map(_p(p), 0, 0, _p(bus), _p(vol))  # Mode A: mix volumes to every bus
    assert n == 2 and np.array_equal(vol[0], p["mix_volumes"][0]) and np.array_equal(vol[1], p["mix_volumes"][0])


# ---- whole path through the world API -----------------------------------------------------------------------------
def _world(orc, **kw):
    cfg = dict(max_instances=4, max_voices=4, max_frames=512, max_spatializers=2, num_buses=3, mix_rate=48000.0)
    cfg.update(kw)
    return orc.OracleMixer(**cfg)


def _emitter(pos, inst=0, area=-1, bus=0, volume_db=0.0):
    e = np.zeros(1, dtype=abi.emitter)
    e["instance"] = inst
    e["area"] = area
    e["bus"] = bus
    e["origin"] = pos
    e["basis_z"] = (0, 0, 1)
    e["volume_db"] = volume_db
    e["max_db"] = 3.0
    e["pitch_scale"] = 1.0
    return e


def test_gain_known_answer_stereo(orc):
    """Source 10 m to the right of an identity listener, inverse model, unit 10, panning 1*0.5:
    multiplier = 1/1.00001, pan g=.25 f=.6 -> (sqrt(.2), sqrt(.8)); filter gain from Q6."""
    with _world(orc) as w:
        w.instance_init([0], 0)
        p = w.gain_compute(_emitter((10, 0, 0)), [abi.identity_listener()])
        m = 1 / 1.00001
        assert np.allclose(p["mix_volumes"][0, 0], (m * math.sqrt(0.2), m * math.sqrt(0.8)), rtol=1e-6)
        assert not p["mix_volumes"][0, 1:].any()
        db_att = (1 - m) * -24.0
        assert abs(p["linear_attenuation"][0] - 10 ** (db_att / 20)) < 1e-6
        assert p["n_bus"][0] == 1 and p["bus"][0, 0] == 0 and p["update_parameters"][0] == 1
        assert np.array_equal(p["bus_volumes"][0, 0], p["mix_volumes"][0])
        assert p["pitch_scale"][0] == 1.0 and p["attenuation_filter_cutoff_hz"][0] == 5000.0


def test_max_distance_and_update_flag(orc):
    """Q5 + Q18 (:361-373, :437-471): out of range => no bus, zero volumes; update_parameters only on
    the first out-of-range tick."""
    with _world(orc) as w:
        w.spatializer_set(0, abi.spatializer_defaults(max_distance=50.0))
        w.instance_init([0], 0)
        L = [abi.identity_listener()]
        p = w.gain_compute(_emitter((25, 0, 0)), L)
        m = (1 / 2.50001) * 0.5  # taper 1 - 25/50
        assert abs(p["mix_volumes"][0, 0, 1] - m * math.sqrt(0.8)) < 1e-6 and p["update_parameters"][0] == 1
        p = w.gain_compute(_emitter((60, 0, 0)), L)
        assert p["n_bus"][0] == 0 and not p["mix_volumes"].any() and p["update_parameters"][0] == 1
        assert p["linear_attenuation"][0] == 0.0  # never set for a skipped listener
        p = w.gain_compute(_emitter((60, 0, 0)), L)
        assert p["update_parameters"][0] == 0
        p = w.gain_compute(_emitter((25, 0, 0)), L)
        assert p["update_parameters"][0] == 1 and p["n_bus"][0] == 1


def test_listeners_max_combine_and_bus_selection(orc):
    """Q7: identical listeners combine by max (idempotent).  Q18: area override bus, reverb bus,
    unknown player bus => Master."""
    with _world(orc) as w:
        w.instance_init([0, 1], 0)
        one = w.gain_compute(_emitter((3, 1, -4)), [abi.identity_listener()])
        two = w.gain_compute(_emitter((3, 1, -4)), [abi.identity_listener(), abi.identity_listener()])
        assert np.array_equal(one["mix_volumes"], two["mix_volumes"])
        area = S.synth.reverb_area(reverb_bus=2, amount=0.5, uniformity=0.0, override_bus=True, bus=1).reshape(1)
        p = w.gain_compute(_emitter((3, 1, -4), inst=1, area=0, bus=0), [abi.identity_listener()], area)
        assert p["n_bus"][0] == 2 and tuple(p["bus"][0, :2]) == (1, 2)
        assert np.allclose(p["bus_volumes"][0, 1], p["mix_volumes"][0] * np.float32(0.5), rtol=1e-7)  # Q17, uniformity 0
        p = w.gain_compute(_emitter((3, 1, -4), inst=1, bus=77), [abi.identity_listener()])
        assert p["n_bus"][0] == 1 and p["bus"][0, 0] == 0


def test_bus_fade_in_steady_and_fade_out(orc):
    """upstream AudioServer ramps: a new proxy fades in from 0, a steady bus keeps
    its volume, a bus that disappears is mixed once more towards 0."""
    F = 512
    with _world(orc) as w:
        w.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=0))
        w.instance_init([0], 0)
        p = np.zeros(1, dtype=abi.params)
        p["mix_volumes"][0, 0] = (0.5, 0.5)
        p["n_bus"] = 2
        p["bus"][0, :2] = (0, 1)
        p["update_parameters"] = 1
        w.params_set([0], p)
        w.instance_start([0])
        w.voice_init([0])
        v = S.synth.make_voices(1)
        src = np.ones((1, F, 2), np.float32)
        bus, _ = w.mix_block(v, src, F)
        t = np.arange(F, dtype=np.float32) / np.float32(F)
        assert np.allclose(bus[0, 0, :, 0], 0.5 * t, atol=1e-7) and np.array_equal(bus[0], bus[1])  # Q15: same volumes to every bus
        assert not bus[2].any()
        bus, _ = w.mix_block(v, src, F)
        assert np.allclose(bus[0, 0, :, 0], 0.5, atol=1e-7)
        p["n_bus"] = 1                                   # bus 1 dropped
        w.params_set([0], p)
        bus, _ = w.mix_block(v, src, F)
        assert np.allclose(bus[0, 0, :, 0], 0.5, atol=1e-7)
        assert np.allclose(bus[1, 0, :, 0], 0.5 * (1 - t), atol=1e-7)
        bus, _ = w.mix_block(v, src, F)
        assert not bus[1].any()
        p["update_parameters"] = 0                       # not pushed to AudioServer (:265)
        p["n_bus"] = 0
        w.params_set([0], p)
        bus, _ = w.mix_block(v, src, F)
        assert np.allclose(bus[0, 0, :, 0], 0.5, atol=1e-7)


def test_mode_a_equals_mode_b_when_steady_and_unfiltered(orc):
    """SURVEY §4 property: filter off, one bus, volumes constant => both modes put vol*x on the bus."""
    sc = S.default_scenario(voices=8, blocks=3, gain_every=100, force_filter_off=True, speaker_mode=abi.SPEAKER_SURROUND_51)
    res = {}
    for mode in (0, 1):
        sc["spat"] = dict(mix_channel_mode=mode)
        with orc.OracleMixer(**S.config_of(sc)) as o:
            res[mode] = S.run(o, sc)["bus"][-1]
    ok, worst, _ = S.sample_close(res[0], res[1], rel=1e-6)
    assert ok, worst


def test_threads_and_shadow_agree(orc):
    sc = S.default_scenario(voices=40, voices_per_instance=2, blocks=2, spat=dict(mix_channel_mode=1), speaker_mode=abi.SPEAKER_SURROUND_71)
    cfg = S.config_of(sc)
    outs = []
    for threads in (1, 4):
        with orc.OracleMixer(**cfg) as o:
            o.spatializer_set(0, S.make_spatializer(sc))
            n_inst = 20
            o.instance_init(np.arange(n_inst), 0)
            p = o.gain_compute(S.synth.make_emitters(n_inst), [abi.identity_listener()])
            p["linear_attenuation"] = 0.0  # unfiltered: the shadow budgets summation order, not the float coefficient ramp
            o.params_set(np.arange(n_inst), p)
            o.instance_start(np.arange(n_inst))
            o.voice_init(np.arange(40))
            src = S.synth.make_sources(40, 512)
            bus, _ = o.mix_block(S.synth.make_voices(40, voices_per_instance=2), src, 512, shadow=(threads == 1), threads=threads)
            outs.append(bus)
            if threads == 1:
                ok, worst, _ = S.sample_close(bus, o.last_bus64)
                assert ok, f"float32 oracle vs float64 shadow: {worst}"
    ok, worst, _ = S.sample_close(outs[0], outs[1])
    assert ok, worst


def test_invalid_arguments_leave_state_untouched(orc):
    with _world(orc) as w:
        for bad in (dict(max_distance=-1.0), dict(emission_angle=91.0), dict(attenuation_model=4), dict(panning_strength=-0.1),
                    dict(doppler_speed_of_sound=0.0)):  # audio_spatializer_3d.cpp:671,696,729,738,759
            with pytest.raises(orc.OracleError):
                w.spatializer_set(0, abi.spatializer_defaults(**bad))
        with pytest.raises(orc.OracleError):
            w.mix_block(S.synth.make_voices(1), np.zeros((1, 511, 2), np.float32), 511)
