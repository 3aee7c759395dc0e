"""GPU tests of the pipelined form (gas_step_device): one launch streams block k and computes gains + plan of block k + 1 on
the control warps of the same kernel.  It runs the same planner / gain code as the one-call-per-block entry points, so its
bus buffers must equal theirs bit for bit where the summation order is the same (one CTA per (class, tile) run) and within
the north-star tolerance otherwise; and it must match the oracle like every other path."""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu
abi, synth = S.abi, S.synth


def _setup(m, V, F, spat, area, em0):
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(**area)], dtype=abi.area)
    m.spatializer_set(0, abi.spatializer_defaults(**spat))
    m.instance_init(inst, 0)
    m.gain_compute(em0, listeners, areas, want_params=False)
    m.instance_start(inst)
    m.voice_init(inst)
    return listeners, areas


CASES = {
    "stream-7.1": dict(V=700, F=512, mode=abi.SPEAKER_SURROUND_71, spat=dict(mix_channel_mode=1, attenuation_filter_db=0.0), peaks=0),
    "stream-stereo-odd": dict(V=333, F=384, mode=abi.SPEAKER_MODE_STEREO, spat=dict(mix_channel_mode=1, attenuation_filter_db=0.0), peaks=0),
    "filter-5.1-modeB": dict(V=200, F=256, mode=abi.SPEAKER_SURROUND_51, spat=dict(mix_channel_mode=1, attenuation_filter_db=-12.0), peaks=3),
    "filter-7.1-modeA": dict(V=150, F=512, mode=abi.SPEAKER_SURROUND_71, spat=dict(mix_channel_mode=0, attenuation_filter_db=-12.0), peaks=0),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("graph", [False, True])
def test_pipelined_steps_match_block_calls_and_oracle(gas, orc, case, graph):
    import torch
    c = CASES[case]
    V, F, mode, blocks = c["V"], c["F"], c["mode"], 6
    C = mode + 1
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=mode, mix_rate=48000.0)
    area = dict(reverb_bus=1, amount=0.5)
    dev = torch.device("cuda", 0)
    voices_h = synth.make_voices(V)
    if c["peaks"]:
        voices_h["flags"][:: c["peaks"]] |= abi.VOICE_WANT_PEAK
    ems_h = [synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.5) for b in range(blocks + 1)]
    src_h = [synth.make_sources(V, F, block=b) for b in range(blocks)]
    voices = torch.from_numpy(voices_h.view(np.uint8).copy()).to(dev)
    ems = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in ems_h]
    src = [torch.from_numpy(s_).to(dev) for s_ in src_h]

    # reference run: gas_gain_compute_device + gas_mix_block_device, block by block
    want_bus, want_peaks = [], []
    with gas.Mixer(**cfg) as m:
        listeners, areas = _setup(m, V, F, c["spat"], area, ems_h[0])
        m.listeners_set(listeners)
        m.areas_set(areas)
        bus = torch.zeros((2, C, F, 2), device=dev)
        peaks = torch.zeros((V, 2), device=dev)
        for b in range(blocks):
            m.gain_compute_device(V, ems[b].data_ptr())
            m.mix_block_device(V, voices.data_ptr(), src[b].data_ptr(), V, F, F, bus.data_ptr(), peaks.data_ptr())
            m.sync()
            want_bus.append(bus.cpu().numpy().copy())
            want_peaks.append(peaks.cpu().numpy().copy())
        want_state = m.voice_state_export(np.arange(V, dtype=np.int32))

    # pipelined run: 4 rotating output buffers
    with gas.Mixer(**cfg) as m:
        listeners, areas = _setup(m, V, F, c["spat"], area, ems_h[0])
        m.listeners_set(listeners)
        m.areas_set(areas)
        busb = [torch.full((2, C, F, 2), 7.0, device=dev) for _ in range(4)]  # poisoned: the step must zero them itself
        peakb = [torch.full((V, 2), 7.0, device=dev) for _ in range(4)]

        def nxt(b):
            return dict(n_emitters=V, d_emitters=ems[b].data_ptr(), n_voices=V, d_voices=voices.data_ptr(), src_rows=V, frames=F,
                        d_bus_out=busb[b % 4].data_ptr(), d_peaks=peakb[b % 4].data_ptr())

        launches0 = m.kernel_launches
        m.step_device(next=nxt(0))  # first call: plans block 0 only
        got_bus, got_peaks = [], []
        if graph:
            # two steps per graph: (stream 0, plan 1), (stream 1, plan 2) ... ; the last step ends the run eagerly
            b = 0
            while b + 2 <= blocks - 1:
                m.capture_begin()
                m.step_device(src[b].data_ptr(), F, next=nxt(b + 1))
                m.step_device(src[b + 1].data_ptr(), F, next=nxt(b + 2))
                g = m.capture_end()
                m.graph_launch(g)
                m.sync()
                m.graph_destroy(g)
                for j in (b, b + 1):
                    got_bus.append(busb[j % 4].cpu().numpy().copy())
                    got_peaks.append(peakb[j % 4].cpu().numpy().copy())
                b += 2
            while b < blocks:
                m.step_device(src[b].data_ptr(), F, next=nxt(b + 1) if b + 1 < blocks else None)
                m.sync()
                got_bus.append(busb[b % 4].cpu().numpy().copy())
                got_peaks.append(peakb[b % 4].cpu().numpy().copy())
                b += 1
        else:
            for b in range(blocks):
                m.step_device(src[b].data_ptr(), F, next=nxt(b + 1) if b + 1 < blocks else None)
            m.sync()
            # all four buffers are live at the end: blocks 2..5; compare those, then rerun for the early ones below
            for b in range(blocks - 4, blocks):
                got_bus.append(busb[b % 4].cpu().numpy().copy())
                got_peaks.append(peakb[b % 4].cpu().numpy().copy())
            want_bus, want_peaks = want_bus[blocks - 4:], want_peaks[blocks - 4:]
        assert m.kernel_launches > launches0
        got_state = m.voice_state_export(np.arange(V, dtype=np.int32))
        # the run is over: the block-call form works again
        m.mix_block_device(V, voices.data_ptr(), src[0].data_ptr(), V, F, F, busb[0].data_ptr())
        m.sync()

    for b, (g_, w_) in enumerate(zip(got_bus, want_bus)):
        assert np.array_equal(S.routing(g_), S.routing(w_)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(g_, w_)
        assert ok, f"block {b}: {nbad} samples differ from the block-call form (worst {worst:.3e})"
    for g_, w_ in zip(got_peaks, want_peaks):
        np.testing.assert_array_equal(g_, w_)
    for name in want_state.dtype.names:
        np.testing.assert_array_equal(got_state[name], want_state[name], err_msg=f"voice state field {name}")


def test_pipelined_steps_match_oracle(gas, orc):
    """The same run against the CPU oracle (gain + mix per block), north-star tolerance."""
    import torch
    V, F, mode, blocks = 512, 512, abi.SPEAKER_SURROUND_71, 4
    C = mode + 1
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=mode, mix_rate=48000.0)
    spat = dict(mix_channel_mode=1, attenuation_filter_db=0.0)
    area = dict(reverb_bus=1, amount=0.5)
    dev = torch.device("cuda", 0)
    voices_h = synth.make_voices(V)
    ems_h = [synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.5) for b in range(blocks)]
    src_h = [synth.make_sources(V, F, block=b) for b in range(blocks)]
    want = []
    with orc.OracleMixer(**cfg) as o:
        listeners, areas = _setup(o, V, F, spat, area, ems_h[0])
        for b in range(blocks):
            o.gain_compute(ems_h[b], listeners, areas, want_params=False)
            bus, _ = o.mix_block(voices_h, src_h[b], F, want_peaks=False)
            want.append(bus)
    voices = torch.from_numpy(voices_h.view(np.uint8).copy()).to(dev)
    ems = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in ems_h]
    src = [torch.from_numpy(s_).to(dev) for s_ in src_h]
    with gas.Mixer(**cfg) as m:
        listeners, areas = _setup(m, V, F, spat, area, ems_h[0])
        m.listeners_set(listeners)
        m.areas_set(areas)
        busb = [torch.zeros((2, C, F, 2), device=dev) for _ in range(blocks)]

        def nxt(b):
            return dict(n_emitters=V, d_emitters=ems[b].data_ptr(), n_voices=V, d_voices=voices.data_ptr(), src_rows=V, frames=F,
                        d_bus_out=busb[b].data_ptr())

        m.step_device(next=nxt(0))
        for b in range(blocks):
            m.step_device(src[b].data_ptr(), F, next=nxt(b + 1) if b + 1 < blocks else None)
        m.sync()
        for b in range(blocks):
            got = busb[b].cpu().numpy()
            assert np.array_equal(S.routing(got), S.routing(want[b])), f"block {b}: routing differs from the oracle"
            ok, worst, nbad = S.sample_close(got, want[b])
            assert ok, f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"


def test_block_calls_refused_while_a_planned_block_waits(gas):
    import torch
    V, F = 64, 128
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=abi.SPEAKER_MODE_STEREO, mix_rate=48000.0)
    dev = torch.device("cuda", 0)
    voices = torch.from_numpy(synth.make_voices(V).view(np.uint8).copy()).to(dev)
    src = torch.zeros((V, F, 2), device=dev)
    bus = [torch.zeros((2, 1, F, 2), device=dev) for _ in range(2)]
    with gas.Mixer(**cfg) as m:
        _setup(m, V, F, dict(mix_channel_mode=1), dict(reverb_bus=1, amount=0.5), synth.make_emitters(V, block=0, dt=F / 48000.0))
        nx = dict(n_voices=V, d_voices=voices.data_ptr(), src_rows=V, frames=F, d_bus_out=bus[0].data_ptr())
        m.step_device(next=nx)
        with pytest.raises(gas.GasError):
            m.mix_block_device(V, voices.data_ptr(), src.data_ptr(), V, F, F, bus[1].data_ptr())
        with pytest.raises(gas.GasError):  # the next block may not reuse the buffers of the block being streamed
            m.step_device(src.data_ptr(), F, next=nx)
        m.step_device(src.data_ptr(), F, next=None)
        m.sync()
        m.mix_block_device(V, voices.data_ptr(), src.data_ptr(), V, F, F, bus[1].data_ptr())
        m.sync()
