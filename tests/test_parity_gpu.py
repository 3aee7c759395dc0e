"""GPU parity: the CUDA path (through the C ABI, libgas_b200.so) against the CPU oracle on identical
seeded inputs.  Tolerance (north star): samples within 1e-5 relative or below -110 dBFS absolute;
bus / pair routing bit-exact; integer fields of the computed parameters bit-exact.
"""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu
abi = S.abi

GAIN_RTOL = 2e-6  # float32 gains: a couple of ulps between CUDA libm and glibc


def _run_both(gas, orc, sc):
    cfg = S.config_of(sc)
    with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
        launches0 = m.kernel_launches
        got = S.run(m, sc)
        assert m.kernel_launches > launches0, "no CUDA kernel was launched"
        want = S.run(o, sc)
    return got, want


def _check(got, want, sc, state=True):
    for b, (pg, pw) in enumerate(zip(got["params"], want["params"])):
        for f in ("update_parameters", "n_bus", "bus"):
            assert np.array_equal(pg[f], pw[f]), f"block {b}: {f} differs"
        for f in ("mix_volumes", "bus_volumes", "pitch_scale", "linear_attenuation", "attenuation_filter_cutoff_hz"):
            np.testing.assert_allclose(pg[f], pw[f], rtol=GAIN_RTOL, atol=1e-9, err_msg=f"block {b}: {f}")
    for b, (bg, bw) in enumerate(zip(got["bus"], want["bus"])):
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"{sc['name']} block {b}: {nbad} samples out of tolerance, worst abs err {worst:.3e}"
    for b, (kg, kw) in enumerate(zip(got["peaks"], want["peaks"])):
        flagged = np.zeros(len(kw), dtype=bool)
        if sc["want_peak_every"]:
            flagged[:: sc["want_peak_every"]] = True
        ok, worst, nbad = S.sample_close(kg[flagged], kw[flagged])
        assert ok, f"block {b}: peaks differ (worst {worst:.3e})"
        assert np.all(kg[~flagged] == 0)
    if state:
        sg, sw = got["state"], want["state"]
        np.testing.assert_allclose(sg["prev_mix_volumes"], sw["prev_mix_volumes"], rtol=GAIN_RTOL, atol=1e-9)
        for f in ("b0", "b1", "b2", "a1", "a2"):
            np.testing.assert_allclose(sg["filter_processors"][f], sw["filter_processors"][f], rtol=1e-4, atol=1e-7)
        for f in ("ha1", "ha2", "hb1", "hb2"):
            ok, worst, nbad = S.sample_close(sg["filter_processors"][f], sw["filter_processors"][f], rel=1e-4)
            assert ok, f"filter history {f}: worst {worst:.3e}"
        ok, worst, nbad = S.sample_close(sg["effect_history"], sw["effect_history"], rel=1e-4)
        assert ok, f"effect history: worst {worst:.3e}"


MODES = {"A": 0, "B": 1}
SPEAKERS = {"stereo": abi.SPEAKER_MODE_STEREO, "3.1": abi.SPEAKER_SURROUND_31, "5.1": abi.SPEAKER_SURROUND_51, "7.1": abi.SPEAKER_SURROUND_71}


@pytest.mark.parametrize("mode", ["A", "B"])
@pytest.mark.parametrize("speakers", ["stereo", "5.1", "7.1"])
def test_streaming_path_filter_off(gas, orc, mode, speakers):
    """K2: no filter (linear_attenuation forced to 0), ramps change every block."""
    sc = S.default_scenario(name=f"stream-{mode}-{speakers}", voices=64, speaker_mode=SPEAKERS[speakers],
                            spat=dict(mix_channel_mode=MODES[mode]), force_filter_off=True, blocks=3)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


@pytest.mark.parametrize("mode", ["A", "B"])
@pytest.mark.parametrize("speakers", ["stereo", "3.1", "5.1", "7.1"])
def test_filter_path(gas, orc, mode, speakers):
    """K3: attenuation high-shelf active (default -24 dB), first block fades the filter in (Q12)."""
    sc = S.default_scenario(name=f"filter-{mode}-{speakers}", voices=48, speaker_mode=SPEAKERS[speakers],
                            spat=dict(mix_channel_mode=MODES[mode]), blocks=3, want_peak_every=3)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


@pytest.mark.parametrize("mode", ["A", "B"])
def test_reverb_area_two_buses(gas, orc, mode):
    """Reverb Area3D: half of the voices send to a second bus; uniformity 0 and > 0."""
    for uni in (0.0, 0.6):
        sc = S.default_scenario(name=f"reverb-{mode}-{uni}", voices=40, speaker_mode=abi.SPEAKER_SURROUND_51,
                                spat=dict(mix_channel_mode=MODES[mode]), area=dict(reverb_bus=1, amount=0.5, uniformity=uni),
                                area_fraction=0.5, force_filter_off=True, blocks=3)
        got, want = _run_both(gas, orc, sc)
        _check(got, want, sc)


def test_attenuation_models_and_max_distance(gas, orc):
    for model in range(4):
        sc = S.default_scenario(name=f"model-{model}", voices=33, speaker_mode=abi.SPEAKER_SURROUND_71,
                                spat=dict(attenuation_model=model, max_distance=100.0, mix_channel_mode=1,
                                          emission_angle_enabled=1, emission_angle=30.0),
                                listeners="two", blocks=2)
        got, want = _run_both(gas, orc, sc)
        _check(got, want, sc)


def test_effect_chain(gas, orc):
    """AudioSpatializerEffect: cascaded non-interpolated biquads before multi-bus sends."""
    for stages in (1, 2, 4):
        chain = [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=stages)]
        sc = S.default_scenario(name=f"effect-{stages}", voices=37, speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=3,
                                effect_chain=chain, effect_gain_binding=0, area=dict(reverb_bus=2, amount=0.4), area_fraction=0.5,
                                blocks=3)
        got, want = _run_both(gas, orc, sc)
        _check(got, want, sc)


def test_polyphony_late_start_silence_and_odd_sizes(gas, orc):
    """Several voices per instance, voices joining at block 2, silent (tail) voices with peaks,
    a voice count that is not a multiple of anything and a non-power-of-two block size."""
    sc = S.default_scenario(name="poly", voices=45, voices_per_instance=3, frames=480, speaker_mode=abi.SPEAKER_SURROUND_51,
                            spat=dict(mix_channel_mode=1), blocks=4, start_late=2, silent_every=7, want_peak_every=5)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)
    sc = S.default_scenario(name="poly-A-off", voices=45, voices_per_instance=3, frames=130, speaker_mode=abi.SPEAKER_MODE_STEREO,
                            spat=dict(mix_channel_mode=0), blocks=4, start_late=2, silent_every=7, force_filter_off=True)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


def test_empty_block_and_invalid_arguments(gas):
    with gas.Mixer(max_voices=8, max_instances=8) as m:
        bus, peaks = m.mix_block(np.zeros(0, dtype=abi.voice), np.zeros((0, 512, 2), np.float32), 512)
        assert bus.shape == (2, 1, 512, 2) and not bus.any()
        v = S.synth.make_voices(2)
        with pytest.raises(gas.GasError):
            m.mix_block(v, np.zeros((2, 511, 2), np.float32), 511)  # odd frame count
        v["voice"][1] = 99
        with pytest.raises(gas.GasError):
            m.mix_block(v, np.zeros((2, 512, 2), np.float32), 512)  # voice slot out of range
        bad = abi.spatializer_defaults(max_distance=-1.0)
        with pytest.raises(gas.GasError):
            m.spatializer_set(0, bad)  # reference audio_spatializer_3d.cpp:671


@pytest.mark.parametrize("tracking", [1, 2])
def test_doppler_tracking(gas, orc, tracking):
    """Doppler pitch (audio_spatializer_3d.cpp:405-434): moving emitters, two moving / rotated listeners."""
    listeners = [S.synth.rotated_listener(velocity=(1.0, -2.0, 0.5)), S.synth.rotated_listener(yaw=-1.1, origin=(-5, 0, 2), velocity=(0, 0, 0))]
    sc = S.default_scenario(name=f"doppler-{tracking}", voices=50, speaker_mode=abi.SPEAKER_SURROUND_51, listeners=listeners, blocks=2,
                            spat=dict(mix_channel_mode=1, doppler_tracking=tracking, doppler_speed_of_sound=200.0))
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)
    assert np.ptp(want["params"][0]["pitch_scale"]) > 1e-3, "Doppler did not move the pitch"


@pytest.mark.parametrize("speakers", ["3.1", "5.1", "7.1"])
@pytest.mark.parametrize("filt", [False, True])
def test_q1_non_integer_tightness_nan_like_the_reference(gas, orc, speakers, filt):
    """un-normalised source direction into SPCAP with a non-integer tightness gives pow(negative, frac) = NaN
    in the reference (audio_spatializer_3d.cpp:391 -> :930).  The CUDA path must produce NaN gains for the same voices and
    NaN on the same (bus, pair, side) outputs — including the pairs the NaN reaches through the masked-out, zero-volume
    proxy sends of Mode B (pinned by oracle/_ref) — and finite values everywhere else."""
    sc = S.default_scenario(name=f"q1-{speakers}-{filt}", voices=64, speaker_mode=SPEAKERS[speakers], blocks=2,
                            spat=dict(mix_channel_mode=1, panning_strength=0.37), force_filter_off=not filt,
                            area=dict(reverb_bus=1, amount=0.5), area_fraction=0.5)
    got, want = _run_both(gas, orc, sc)
    assert np.isnan(want["params"][0]["mix_volumes"]).any(), "scenario did not reach the NaN case"
    for b in range(len(want["bus"])):
        assert np.array_equal(np.isnan(got["bus"][b]), np.isnan(want["bus"][b])), f"block {b}: NaN pattern differs"
    _check(got, want, sc, state=False)
    # Mode A keeps the NaN on the pairs whose own mix volume is NaN (one proxy, no masked-out sends)
    sc = S.default_scenario(name=f"q1-A-{speakers}", voices=64, speaker_mode=SPEAKERS[speakers], blocks=2,
                            spat=dict(mix_channel_mode=0, panning_strength=1.3), force_filter_off=not filt)
    got, want = _run_both(gas, orc, sc)
    for b in range(len(want["bus"])):
        assert np.array_equal(np.isnan(got["bus"][b]), np.isnan(want["bus"][b])), f"Mode A block {b}: NaN pattern differs"
    _check(got, want, sc, state=False)


def test_scaled_send_classes(gas, orc):
    """Reverb sends with uniformity 0 are the direct send times the area's amount: the streaming kernel keeps one row
    group for such voices and scales at the flush.  Several areas' worth of amounts, a third (override) bus, and the
    block in which the sends fade in (not scaled: quadratic ramp) all have to match the oracle."""
    for amount, override in ((0.5, False), (0.3, True), (1.7, False)):
        sc = S.default_scenario(name=f"scaled-{amount}", voices=96, speaker_mode=abi.SPEAKER_SURROUND_71, num_buses=3,
                                spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0), force_filter_off=True,
                                area=dict(reverb_bus=1, amount=amount, override_bus=override, bus=2), area_fraction=0.5, blocks=4)
        got, want = _run_both(gas, orc, sc)
        _check(got, want, sc)


def test_full_size_configs_1_and_3(gas, orc):
    """BASELINE.json configs[1] (1024 voices, 5.1, inverse-square + attenuation filter) and configs[3] (4096-voice
    AudioSpatializerEffect chain before three buses) at FULL size against the oracle."""
    sc = S.default_scenario(name="cfg1-full", voices=1024, speaker_mode=abi.SPEAKER_SURROUND_51, blocks=2,
                            spat=dict(mix_channel_mode=1, attenuation_model=abi.ATTENUATION_INVERSE_SQUARE_DISTANCE),
                            area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, want_peak_every=16)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)
    chain = [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=2),
             dict(mode=abi.FILTER_LOWPASS, cutoff_hz=9000.0, resonance=0.7, gain=1.0, stages=1)]
    sc = S.default_scenario(name="cfg3-full", voices=4096, speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=3, effect_chain=chain,
                            effect_gain_binding=0, area=dict(reverb_bus=2, amount=0.4, override_bus=True, bus=1), area_fraction=0.5, blocks=2)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


def test_class_table_overflow_falls_back_to_the_generic_class(gas, orc):
    """More distinct routing classes than the plan has slots: the voices whose class found no slot go through the generic
    class of the voice-parallel kernel and are still mixed (round 1 dropped them).  150+ classes are made with one
    instance per distinct (bus mask, scale) combination through gas_params_set."""
    V, F = 200, 128
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=16, speaker_mode=abi.SPEAKER_MODE_STEREO,
               mix_rate=48000.0)
    rng = np.random.default_rng(5)
    inst = np.arange(V, dtype=np.int32)
    p = np.zeros(V, dtype=abi.params)
    p["pitch_scale"] = 1.0
    p["update_parameters"] = 1
    p["mix_volumes"][:, 0, :] = rng.uniform(0.2, 1.0, (V, 2)).astype(np.float32)
    for i in range(V):  # instance i: Master plus one more bus at its own scale: 200 distinct (mask, scale) classes > 122 slots
        p["n_bus"][i] = 2
        p["bus"][i, :2] = [0, 1 + i % 15]
        p["bus_volumes"][i, 0, 0] = p["mix_volumes"][i, 0]
        p["bus_volumes"][i, 1, 0] = p["mix_volumes"][i, 0] * np.float32(0.2 + 0.003 * i)
    voices = S.synth.make_voices(V)
    src = S.synth.make_sources(V, F)
    out = []
    for mk in (lambda: gas.Mixer(**cfg), lambda: orc.OracleMixer(**cfg)):
        with mk() as m:
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
            m.instance_init(inst, 0)
            m.params_set(inst, p)
            m.instance_start(inst)
            m.voice_init(inst)
            blocks = []
            for b in range(3):
                m.params_set(inst, p)
                blocks.append(m.mix_block(voices, src, F, want_peaks=False)[0])
            out.append(blocks)
    for b, (bg, bw) in enumerate(zip(*out)):
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"


def test_filter_tile_long_effect_chains(gas, orc):
    """Filter-tile path of the voice-parallel kernel (gas_mix_voice.cu): chains of up to four biquads per side run out of
    registers, longer ones (here 2 x 4 and 4 + 3 + 1 stages) out of local memory; both against the oracle, with peaks."""
    for name, chain in (
            ("3fx-fast", [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=1),
                          dict(mode=abi.FILTER_LOWPASS, cutoff_hz=9000.0, resonance=0.7, gain=1.0, stages=2),
                          dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=2500.0, resonance=0.9, gain=1.6, stages=1)]),
            ("2x4-slow", [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.5, stages=4),
                          dict(mode=abi.FILTER_LOWPASS, cutoff_hz=11000.0, resonance=0.7, gain=1.0, stages=4)]),
            ("431-slow", [dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=3000.0, resonance=1.0, gain=0.7, stages=4),
                          dict(mode=abi.FILTER_LOWPASS, cutoff_hz=12000.0, resonance=0.7, gain=1.0, stages=3),
                          dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=6000.0, resonance=0.8, gain=1.2, stages=1)])):
        sc = S.default_scenario(name=f"ft-{name}", voices=53, frames=200, speaker_mode=abi.SPEAKER_SURROUND_51, num_buses=3,
                                effect_chain=chain, effect_gain_binding=0, area=dict(reverb_bus=2, amount=0.4), area_fraction=0.5,
                                blocks=3, want_peak_every=4)
        got, want = _run_both(gas, orc, sc)
        _check(got, want, sc)


def test_filter_tile_many_units_ragged_frames(gas, orc):
    """Filter-tile path, Mode A with the attenuation filter: enough voices that a CTA works through several units, a frame
    count that is not a multiple of the 64-frame tile, silent rows, late starters and peaks."""
    sc = S.default_scenario(name="ft-A-many", voices=1100, frames=200, speaker_mode=abi.SPEAKER_SURROUND_71,
                            spat=dict(mix_channel_mode=0), area=dict(reverb_bus=1, amount=0.5, uniformity=0.4), area_fraction=0.3,
                            blocks=3, start_late=1, silent_every=9, want_peak_every=6)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


def test_mode_a_filter_twelve_sends_generic_class(gas, orc):
    """Mode A voices with the filter on whose six buses all change between blocks: six sends fade out while six fade in.  More
    than two buses on a side take the generic class of the voice-parallel kernel (one voice per unit, sends two at a time, six
    passes from the same filter state), next to ordinary classes on the filter-tile path in the same launch."""
    V, F = 70, 256
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=16, speaker_mode=abi.SPEAKER_SURROUND_71,
               mix_rate=48000.0)
    rng = np.random.default_rng(11)
    inst = np.arange(V, dtype=np.int32)

    def params(block):
        p = np.zeros(V, dtype=abi.params)
        p["pitch_scale"] = 1.0
        p["update_parameters"] = 1
        p["linear_attenuation"] = 0.4          # >= 0.001: the high-shelf runs (audio_spatializer_3d.cpp:503)
        p["attenuation_filter_cutoff_hz"] = 5000.0
        p["mix_volumes"] = rng.uniform(0.2, 1.0, (V, 4, 2)).astype(np.float32)
        p["n_bus"] = 6
        for i in range(V):
            first = (1 + 6 * (block % 2)) if i % 2 == 0 else 1 + (block + i) % 9
            p["bus"][i] = np.arange(first, first + 6)
            p["bus_volumes"][i] = rng.uniform(0.1, 0.9, (6, 4, 2)).astype(np.float32)
        return p

    ps = [params(b) for b in range(4)]
    voices = S.synth.make_voices(V)
    srcs = [S.synth.make_sources(V, F, block=b) for b in range(4)]
    out = []
    for mk in (lambda: gas.Mixer(**cfg), lambda: orc.OracleMixer(**cfg)):
        with mk() as m:
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=0))
            m.instance_init(inst, 0)
            m.params_set(inst, ps[0])
            m.instance_start(inst)
            m.voice_init(inst)
            blocks = []
            for b in range(4):
                m.params_set(inst, ps[b])
                blocks.append(m.mix_block(voices, srcs[b], F, want_peaks=False)[0])
            out.append(blocks)
    for b, (bg, bw) in enumerate(zip(*out)):
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"
    assert (S.routing(out[1][1]).reshape(16, -1).any(axis=1).sum()) >= 12, "scenario did not reach twelve sends"


@pytest.mark.parametrize("mode", ["A", "B"])
def test_q1_nan_gains_of_a_silent_voice_still_reach_the_buses(gas, orc, mode):
    """SURVEY Q1 meets the silent tail: a voice without source frames (the reference's zero-filled playback_buffer,
    audio_spatializer.cpp:405-408) whose pan gains are NaN still poisons its buses, 0 * NaN being NaN.  Found by
    tools/fuzz_parity.py (seed 7, case 196): the planner used to drop every silent unfiltered voice."""
    sc = S.default_scenario(name=f"q1-silent-{mode}", voices=64, speaker_mode=abi.SPEAKER_SURROUND_51, blocks=2, silent_every=2,
                            spat=dict(mix_channel_mode=MODES[mode], panning_strength=1.5), force_filter_off=True,
                            area=dict(reverb_bus=1, amount=0.3, uniformity=1.0), area_fraction=0.5)
    got, want = _run_both(gas, orc, sc)
    silent = (S.synth.make_voices(64)["voice"] % 2) == 1
    nan_voice = np.isnan(want["params"][0]["mix_volumes"]).any(axis=(1, 2))
    assert (nan_voice & silent).any(), "scenario has no silent voice with NaN gains"
    for b in range(len(want["bus"])):
        assert np.array_equal(np.isnan(got["bus"][b]), np.isnan(want["bus"][b])), f"block {b}: NaN pattern differs"
    _check(got, want, sc, state=False)


@pytest.mark.parametrize("speakers", ["stereo", "3.1", "5.1", "7.1"])
def test_filter_tile_mode_b_many_units_ragged_frames(gas, orc, speakers):
    """Filter-tile path, Mode B with the attenuation filter (2C biquads per voice): several units per CTA, a frame count that is
    not a multiple of the 64-frame tile, two buses, silent rows, late starters and peaks (max over the voice's pairs)."""
    sc = S.default_scenario(name=f"ft-B-many-{speakers}", voices=420, frames=200, speaker_mode=SPEAKERS[speakers],
                            spat=dict(mix_channel_mode=1), area=dict(reverb_bus=1, amount=0.5, uniformity=0.4), area_fraction=0.3,
                            blocks=3, start_late=1, silent_every=9, want_peak_every=5)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)


def test_mixed_spatializers_in_one_block(gas, orc):
    """Four AudioSpatializer resources in one context — 3D Mode A and Mode B with the attenuation filter, an effect chain, and an
    unfiltered Mode B one (streaming kernel) — instances dealt round-robin to them: every kind of class meets in the same launch of the
    voice-parallel kernel (filter-tile units of three different forms dealt to the same CTAs) beside the streaming kernel."""
    V, F, blocks = 333, 320, 3
    mode = abi.SPEAKER_SURROUND_51
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=4, num_buses=3, speaker_mode=mode, mix_rate=48000.0)
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener(), S.synth.rotated_listener()], dtype=abi.listener)
    areas = np.array([S.synth.reverb_area(n_listeners=2, reverb_bus=2, amount=0.4, uniformity=0.5)], dtype=abi.area)
    fx = S.make_spatializer(S.default_scenario(effect_chain=[dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=3500.0, resonance=1.0, gain=0.4, stages=2),
                                                             dict(mode=abi.FILTER_LOWPASS, cutoff_hz=9000.0, resonance=0.7, gain=1.0, stages=1)],
                                               effect_gain_binding=0))
    spats = [abi.spatializer_defaults(mix_channel_mode=0), abi.spatializer_defaults(mix_channel_mode=1), fx,
             abi.spatializer_defaults(mix_channel_mode=1, attenuation_filter_db=0.0)]
    voices = S.synth.make_voices(V)
    voices["flags"][::7] |= abi.VOICE_WANT_PEAK
    out = []
    for mk in (lambda: gas.Mixer(**cfg), lambda: orc.OracleMixer(**cfg)):
        with mk() as m:
            for k, sp in enumerate(spats):
                m.spatializer_set(k, sp)
            m.instance_init(inst, inst % len(spats))
            res = []
            for b in range(blocks):
                em = S.synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.4)
                m.gain_compute(em, listeners, areas, want_params=False)
                if b == 0:
                    m.instance_start(inst)
                    m.voice_init(inst)
                res.append(m.mix_block(voices, S.synth.make_sources(V, F, block=b), F, want_peaks=True))
            out.append(res)
    for b, ((bg, pg), (bw, pw)) in enumerate(zip(*out)):
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"
        ok, worst, nbad = S.sample_close(pg[::7], pw[::7])
        assert ok, f"block {b}: peaks differ (worst {worst:.3e})"


@pytest.mark.parametrize("kind", ["A", "B", "E"])
def test_filter_tile_without_a_bus_tile(gas, orc, kind):
    """2048-frame blocks on three 7.1 buses: the bus layout (192 KiB) does not fit the CTA's shared-memory tile, so the voice-parallel
    kernel adds its contraction results straight into the bus buffers."""
    kw = dict(A=dict(spat=dict(mix_channel_mode=0)), B=dict(spat=dict(mix_channel_mode=1)),
              E=dict(effect_chain=[dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=2)],
                     effect_gain_binding=0))[kind]
    sc = S.default_scenario(name=f"ft-notile-{kind}", voices=70, frames=2048, speaker_mode=abi.SPEAKER_SURROUND_71, num_buses=3, blocks=2,
                            area=dict(reverb_bus=1, amount=0.5, uniformity=0.3), area_fraction=0.5, want_peak_every=3, **kw)
    got, want = _run_both(gas, orc, sc)
    _check(got, want, sc)
