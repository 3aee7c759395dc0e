"""GPU tests of the rest of the C-ABI surface: CUDA-graph capture of the device-resident calls, per-kernel timing,
voice-state export / import (checkpoint, re-sharding), run-time AudioServer globals, class-table overflow."""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu
abi, synth = S.abi, S.synth


def _scene(m, V, F, mode, spat=None, area_fraction=0.5):
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5)], dtype=abi.area)
    m.spatializer_set(0, abi.spatializer_defaults(**(spat or dict(mix_channel_mode=1))))
    m.instance_init(inst, 0)
    m.gain_compute(synth.make_emitters(V, block=0, dt=F / 48000.0, area_fraction=area_fraction), listeners, areas, want_params=False)
    m.instance_start(inst)
    m.voice_init(inst)
    return listeners, areas


def test_graph_replay_matches_eager_calls(gas):
    """gas_capture_begin/end + gas_graph_launch of (mix block, gain for the next block) equals the same calls made eagerly."""
    import torch
    V, F, mode, blocks = 256, 256, abi.SPEAKER_SURROUND_51, 4
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=mode, mix_rate=48000.0)
    dev = torch.device("cuda", 0)
    voices = torch.from_numpy(synth.make_voices(V).view(np.uint8).copy()).to(dev)
    src = [torch.from_numpy(synth.make_sources(V, F, block=b)).to(dev) for b in range(blocks)]
    ems = [torch.from_numpy(synth.make_emitters(V, block=b + 1, dt=F / 48000.0, area_fraction=0.5).view(np.uint8).copy()).to(dev)
           for b in range(blocks)]
    outs = []
    for use_graph in (False, True):
        with gas.Mixer(**cfg) as m:
            listeners, areas = _scene(m, V, F, mode, spat=dict(mix_channel_mode=1, attenuation_filter_db=-6.0))
            m.listeners_set(listeners)
            m.areas_set(areas)
            bus = torch.zeros((2, mode + 1, F, 2), device=dev)
            got = []
            for b in range(blocks):
                def step():
                    m.mix_block_device(V, voices.data_ptr(), src[b].data_ptr(), V, F, F, bus.data_ptr())
                    m.gain_compute_device(V, ems[b].data_ptr())
                if use_graph:
                    m.capture_begin()
                    step()
                    g = m.capture_end()
                    launches0 = m.kernel_launches
                    m.graph_launch(g)
                    assert m.kernel_launches > launches0
                    m.sync()
                    m.graph_destroy(g)
                else:
                    step()
                    m.sync()
                got.append(bus.cpu().numpy().copy())
            outs.append(got)
    for e, g in zip(*outs):
        assert e.any()
        ok, worst, nbad = S.sample_close(g, e)
        assert ok, f"graph replay differs from eager calls: {nbad} samples, worst {worst:.3e}"


def test_voice_kernel_beside_stream_kernel_matches_default_order(gas, monkeypatch):
    """GAS_K3_PARALLEL=1 (K3 on its own stream beside K2, K2 adding straight into the bus buffers) gives the same block
    as the default prologue -> K2 -> K3 order, eagerly and in a replayed graph; filtered and unfiltered voices mixed."""
    import torch
    V, F, mode, blocks = 512, 256, abi.SPEAKER_SURROUND_71, 3
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=mode, mix_rate=48000.0)
    dev = torch.device("cuda", 0)
    voices = torch.from_numpy(synth.make_voices(V).view(np.uint8).copy()).to(dev)
    src = [torch.from_numpy(synth.make_sources(V, F, block=b)).to(dev) for b in range(blocks)]
    ems = [torch.from_numpy(synth.make_emitters(V, block=b + 1, dt=F / 48000.0, area_fraction=0.5).view(np.uint8).copy()).to(dev)
           for b in range(blocks)]
    outs = []
    for parallel, use_graph in ((False, False), (True, False), (True, True)):
        if parallel:
            monkeypatch.setenv("GAS_K3_PARALLEL", "1")
        else:
            monkeypatch.delenv("GAS_K3_PARALLEL", raising=False)
        with gas.Mixer(**cfg) as m:
            # -6 dB filter attenuation: near voices keep the attenuation filter (K3), far ones fall below 0.001 and stream (K2)
            listeners, areas = _scene(m, V, F, mode, spat=dict(mix_channel_mode=1, attenuation_filter_db=-6.0))
            m.listeners_set(listeners)
            m.areas_set(areas)
            bus = torch.zeros((2, mode + 1, F, 2), device=dev)
            got = []
            for b in range(blocks):
                def step():
                    m.mix_block_device(V, voices.data_ptr(), src[b].data_ptr(), V, F, F, bus.data_ptr())
                    m.gain_compute_device(V, ems[b].data_ptr())
                if use_graph:
                    m.capture_begin()
                    step()
                    g = m.capture_end()
                    m.graph_launch(g)
                    m.sync()
                    m.graph_destroy(g)
                else:
                    step()
                    m.sync()
                got.append(bus.cpu().numpy().copy())
            outs.append(got)
    monkeypatch.delenv("GAS_K3_PARALLEL", raising=False)
    for other in outs[1:]:
        for e, g in zip(outs[0], other):
            assert e.any()
            ok, worst, nbad = S.sample_close(g, e)
            assert ok, f"parallel K3 differs from the default order: {nbad} samples, worst {worst:.3e}"


def test_profile_counts_every_kernel(gas):
    V, F = 128, 256
    with gas.Mixer(max_instances=V, max_voices=V, max_frames=F, speaker_mode=abi.SPEAKER_MODE_STEREO) as m:
        _scene(m, V, F, abi.SPEAKER_MODE_STEREO)
        src = synth.make_sources(V, F)
        m.profile_enable(True)
        for _ in range(3):
            m.mix_block(synth.make_voices(V), src, F, want_peaks=False)
        prof = m.profile_read()
        m.profile_enable(False)
    for kind in ("prologue", "mix_stream", "mix_voice", "none"):
        ms, n = prof[kind]
        assert n == 3 and ms > 0.0, f"{kind}: {n} launches, {ms} ms"
    assert prof["none"][0] < prof["mix_stream"][0]  # the empty pair (the timer's own reading) is the smallest of them


def test_voice_state_export_import_resumes_bit_identically(gas):
    """Checkpoint / re-sharding: a second context that imports the voice state (and replays the parameters) continues
    the filtered mix exactly where the first one stopped."""
    V, F, mode = 64, 256, abi.SPEAKER_SURROUND_31
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=mode, mix_rate=48000.0)
    voices = synth.make_voices(V)
    ids = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5)], dtype=abi.area)
    ems = [synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.5) for b in range(4)]
    with gas.Mixer(**cfg) as a:
        _scene(a, V, F, mode)  # default filter: active for every voice
        for b in range(2):
            a.gain_compute(ems[b], listeners, areas, want_params=False)
            a.mix_block(voices, synth.make_sources(V, F, block=b), F, want_peaks=False)
        state = a.voice_state_export(ids)
        params = a.params_get(ids)
        assert np.abs(state["filter_processors"]["ha1"]).max() > 0
        with gas.Mixer(**cfg) as b_:
            b_.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
            b_.instance_init(ids, 0)
            b_.params_set(ids, params)
            b_.instance_start(ids)
            b_.voice_state_import(ids, state)
            # one block on both to get b_'s bus details past their fade-in, compared from the next block on
            outs = []
            for m in (a, b_):
                got = []
                for blk in (2, 3):
                    m.gain_compute(ems[blk], listeners, areas, want_params=False)
                    got.append(m.mix_block(voices, synth.make_sources(V, F, block=blk), F, want_peaks=False)[0])
                outs.append(got)
            sa, sb = a.voice_state_export(ids), b_.voice_state_export(ids)
    ok, worst, nbad = S.sample_close(outs[1][1], outs[0][1])
    assert ok, f"resumed context diverges: {nbad} samples, worst {worst:.3e}"
    np.testing.assert_allclose(sb["prev_mix_volumes"], sa["prev_mix_volumes"], rtol=0, atol=0)
    ok, worst, nbad = S.sample_close(sb["filter_processors"]["ha1"], sa["filter_processors"]["ha1"], rel=1e-4)
    assert ok, f"filter history diverges: worst {worst:.3e}"


def test_speaker_mode_and_mix_rate_change_at_run_time(gas, orc):
    """AudioServer::get_speaker_mode / get_mix_rate are read per tick (audio_spatializer_3d.cpp:59,113,506): changing them
    between blocks behaves like a context created with the new values."""
    V, F = 48, 256
    sc = S.default_scenario(voices=V, frames=F, speaker_mode=abi.SPEAKER_SURROUND_71, spat=dict(mix_channel_mode=1), blocks=2, mix_rate=44100.0)
    cfg = S.config_of(sc)
    start = dict(cfg, speaker_mode=abi.SPEAKER_MODE_STEREO, mix_rate=48000.0)
    with gas.Mixer(**start) as m, orc.OracleMixer(**cfg) as o:
        m.set_speaker_mode(abi.SPEAKER_SURROUND_71)
        m.set_mix_rate(44100.0)
        assert m.channels == 4
        got, want = S.run(m, sc, collect_state=False), S.run(o, sc, collect_state=False)
    for bg, bw in zip(got["bus"], want["bus"]):
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok and np.array_equal(S.routing(bg), S.routing(bw)), f"{nbad} samples, worst {worst:.3e}"


def test_many_routing_classes_match_the_oracle(gas, orc):
    """78 distinct streaming classes (12 buses: 66 two-bus combinations + 12 single-bus ones), several voices each, two blocks
    (the second one ramps from the first one's volumes): more classes than the streaming kernel's one-class-per-lane partition
    handles, so its fallback path is what gets compared with the oracle here."""
    B, F, per = 12, 128, 5
    combos = [(a, b) for a in range(B) for b in range(a + 1, B)] + [(a, a) for a in range(B)]
    V = len(combos) * per
    ids = np.arange(V, dtype=np.int32)
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=B, speaker_mode=abi.SPEAKER_MODE_STEREO, mix_rate=48000.0)
    rng = np.random.default_rng(7)

    def params(block):
        p = np.zeros(V, dtype=abi.params)
        p["mix_volumes"][:, 0, :] = 0.25 + 0.5 * rng.random((V, 2)).astype(np.float32)
        p["pitch_scale"], p["update_parameters"] = 1.0, 1
        for k in range(V):
            a, b = combos[k // per]
            p["n_bus"][k] = 1 if a == b else 2
            p["bus"][k, 0], p["bus"][k, 1] = a, b
            p["bus_volumes"][k, 0, 0, :] = 0.5 - 0.1 * block
            p["bus_volumes"][k, 1, 0, :] = 0.25 + 0.1 * block
        return p

    ps = [params(b) for b in range(2)]
    voices = synth.make_voices(V)
    outs = []
    for make in (lambda: gas.Mixer(**cfg), lambda: orc.OracleMixer(**cfg)):
        with make() as m:
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
            m.instance_init(ids, 0)
            m.params_set(ids, ps[0])
            m.instance_start(ids)
            m.voice_init(ids)
            got = []
            for b in range(2):
                if b:
                    m.params_set(ids, ps[b])
                got.append(m.mix_block(voices, synth.make_sources(V, F, block=b), F, want_peaks=False)[0])
            outs.append(got)
    for g, e in zip(*outs):
        assert e.any()
        ok, worst, nbad = S.sample_close(g, e)
        assert ok, f"{nbad} samples differ from the oracle, worst {worst:.3e}"


def test_class_table_overflow_is_reported_and_mixed(gas, orc):
    """More distinct routing classes than the plan has dynamic slots (122): the voices concerned go through the generic
    class of the voice-parallel kernel — the mix stays complete (round 1 dropped them) — and gas_status_flags says so.
    16 buses give 120 two-bus combinations + 16 single-bus ones."""
    B, F = 16, 64
    combos = [(a, b) for a in range(B) for b in range(a + 1, B)] + [(a, a) for a in range(B)]
    V = len(combos)
    ids = np.arange(V, dtype=np.int32)
    p = np.zeros(V, dtype=abi.params)
    p["mix_volumes"][:, 0, :] = 0.5
    p["pitch_scale"], p["update_parameters"] = 1.0, 1
    for k, (a, b) in enumerate(combos):
        p["n_bus"][k] = 1 if a == b else 2
        p["bus"][k, 0], p["bus"][k, 1] = a, b
        p["bus_volumes"][k, 0, 0, :] = 0.5
        p["bus_volumes"][k, 1, 0, :] = (0.25, 0.125)  # not a multiple of send 0: no scaled class
    voices, src = synth.make_voices(V), synth.make_sources(V, F)
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=B, speaker_mode=abi.SPEAKER_MODE_STEREO)
    out = []
    flags = 0
    for mk in (lambda: gas.Mixer(**cfg), lambda: orc.OracleMixer(**cfg)):
        with mk() as m:
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
            m.instance_init(ids, 0)
            m.params_set(ids, p)
            m.instance_start(ids)
            m.voice_init(ids)
            out.append([m.mix_block(voices, src, F, want_peaks=False)[0] for _ in range(3)])
            if hasattr(m, "status_flags"):
                flags = m.status_flags()
    assert flags & abi.STATUS_CLASS_OVERFLOW, "the overflow was not reported"
    for b, (bg, bw) in enumerate(zip(*out)):
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst {worst:.3e}"


def test_class_slots_are_recycled(gas):
    """A class slot that stayed empty for GAS_CLS_IDLE_BLOCKS blocks is handed back: a long session that walks through more
    routing classes than there are slots, a few at a time, never overflows."""
    B, F, V = 16, 64, 8
    ids = np.arange(V, dtype=np.int32)
    voices, src = synth.make_voices(V), synth.make_sources(V, F)
    with gas.Mixer(max_instances=V, max_voices=V, max_frames=F, num_buses=B, speaker_mode=abi.SPEAKER_MODE_STEREO) as m:
        m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
        m.instance_init(ids, 0)
        p = np.zeros(V, dtype=abi.params)
        p["mix_volumes"][:, 0, :] = 0.5
        p["pitch_scale"], p["update_parameters"], p["n_bus"] = 1.0, 1, 2
        p["bus_volumes"][:, 0, 0, :] = 0.5
        p["bus_volumes"][:, 1, 0, :] = (0.25, 0.125)
        started = False
        for rnd in range(40):  # 40 rounds x 8 new bus pairs = 320 classes over the session (plus their fade-in / fade-out variants)
            for k in range(V):
                n = rnd * V + k
                p["bus"][k, 0], p["bus"][k, 1] = n % 15, 15 - (n // 15) % 15 if (n % 15) != 15 - (n // 15) % 15 else (n + 1) % 15
            m.params_set(ids, p)
            if not started:
                m.instance_start(ids)
                m.voice_init(ids)
                started = True
            for _ in range(10):
                bus, _ = m.mix_block(voices, src, F, want_peaks=False)
                assert np.isfinite(bus).all()
        assert not (m.status_flags() & abi.STATUS_CLASS_OVERFLOW), "slots were not recycled"


@pytest.mark.parametrize("gain", [0.0, 0.0009, 0.25, 1.0])
def test_per_call_process_frames_and_mix_channel(gas, orc, gain):
    """gas_process_frames / gas_mix_channel: the reference's per-call virtuals (audio_spatializer.h:146,148) on one voice,
    against the oracle's twins (which tests/test_oracle_vs_ref.py pins bit for bit to the reference's own functions):
    odd frame counts, three state-carrying blocks, filter threshold and first-block coefficient fade-in."""
    import ctypes as C
    lo = orc.load()
    rng = np.random.default_rng(31)
    F = 255  # the per-call entry points take any frame count, like the reference
    p = np.zeros(1, dtype=abi.params)
    p["pitch_scale"], p["update_parameters"], p["n_bus"] = 1.0, 1, 1
    p["attenuation_filter_cutoff_hz"] = 4000.0
    for mode_b in (0, 1):
        with gas.Mixer(max_instances=2, max_voices=2, max_frames=256, num_buses=2, speaker_mode=abi.SPEAKER_SURROUND_71, mix_rate=48000.0) as m:
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=mode_b))
            m.instance_init([0], 0)
            m.voice_init([1])
            st = np.zeros(1, dtype=abi.voice_state)
            for blk in range(3):
                p["mix_volumes"] = rng.uniform(0, 1, (1, 4, 2)).astype(np.float32)
                p["bus_volumes"][0, 0] = p["mix_volumes"][0]
                p["linear_attenuation"] = gain
                m.params_set([0], p)
                src = rng.uniform(-0.5, 0.5, (F, 2)).astype(np.float32)
                if mode_b:
                    for ch in range(4):
                        want = np.zeros((F, 2), np.float32)
                        lo.orc_mix_channel_3d(C.c_void_p(p.ctypes.data), C.c_void_p(st.ctypes.data), 48000.0, ch, C.c_void_p(want.ctypes.data),
                                              C.c_void_p(src.ctypes.data), F)
                        got = m.mix_channel(0, 1, ch, src)
                        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=f"mix_channel block {blk} pair {ch}")
                else:
                    want = np.zeros((F, 2), np.float32)
                    lo.orc_process_frames_3d(C.c_void_p(p.ctypes.data), C.c_void_p(st.ctypes.data), 48000.0, C.c_void_p(want.ctypes.data),
                                             C.c_void_p(src.ctypes.data), F)
                    got = m.process_frames(0, 1, src)
                    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=f"process_frames block {blk}")
                dev = m.voice_state_export([1])
                np.testing.assert_allclose(dev["prev_mix_volumes"], st["prev_mix_volumes"], rtol=0, atol=0)
            with pytest.raises(gas.GasError):
                m.mix_channel(0, 1, 4, src)  # ERR_FAIL_INDEX of get_filter_processor (audio_spatializer_3d.cpp:888)


def test_per_call_process_frames_effect_chain(gas, orc):
    import ctypes as C
    lo = orc.load()
    rng = np.random.default_rng(37)
    F = 200
    chain = np.zeros(1, dtype=abi.effect_chain)
    chain["n_effects"] = 2
    chain["effects"][0, 0] = (abi.FILTER_HIGHSHELF, 4000.0, 1.0, 0.3, 2)
    chain["effects"][0, 1] = (abi.FILTER_LOWPASS, 9000.0, 0.7, 1.0, 1)
    spat = abi.spatializer_defaults()
    spat["kind"] = abi.SPATIALIZER_EFFECT
    spat["chain"] = chain[0]
    with gas.Mixer(max_instances=1, max_voices=1, max_frames=256, num_buses=2, mix_rate=44100.0) as m:
        m.spatializer_set(0, spat)
        m.instance_init([0], 0)
        m.voice_init([0])
        st = np.zeros(1, dtype=abi.voice_state)
        for blk in range(3):
            src = rng.uniform(-0.5, 0.5, (F, 2)).astype(np.float32)
            want = np.zeros((F, 2), np.float32)
            lo.orc_process_frames_effect(C.c_void_p(chain.ctypes.data), C.c_void_p(st.ctypes.data), 44100.0, C.c_void_p(want.ctypes.data),
                                         C.c_void_p(src.ctypes.data), F)
            got = m.process_frames(0, 0, src)
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=f"effect chain block {blk}")
