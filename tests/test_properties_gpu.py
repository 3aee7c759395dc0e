"""GPU tests at BASELINE.json's full sizes through size-independent properties of the mix (the oracle would take
minutes there), and oracle parity on corners of the stress sweep (configs[4]: 256-65536 voices x 128-2048-frame
blocks, stereo to 7.1).  Everything goes through the C ABI."""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu
abi, synth = S.abi, S.synth

V_FULL, F_FULL = 16384, 512


def _setup(m, V, F, speaker_mode, area_fraction=0.25, spat=None, block=0):
    spat = spat or dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0)
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
    m.spatializer_set(0, abi.spatializer_defaults(**spat))
    m.instance_init(inst, 0)
    em = synth.make_emitters(V, block=block, dt=F / 48000.0, area_fraction=area_fraction, r_min=10.0, r_max=120.0)
    m.gain_compute(em, listeners, areas, want_params=False)
    m.instance_start(inst)
    m.voice_init(inst)
    return listeners, areas


def _cfg(V, F, speaker_mode, num_buses=2):
    return dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=num_buses, speaker_mode=speaker_mode,
                mix_rate=48000.0)


def _blocks(m, V, F, src, n=2, listeners=None, areas=None, voices=None):
    voices = synth.make_voices(V) if voices is None else voices
    out = []
    for b in range(n):
        if b:
            em = synth.make_emitters(V, block=b, dt=F / 48000.0, area_fraction=0.25, r_min=10.0, r_max=120.0)
            m.gain_compute(em, listeners, areas, want_params=False)
        bus, _ = m.mix_block(voices, src, F, want_peaks=False)
        out.append(bus)
    return out


@pytest.fixture(scope="module")
def full_src():
    rng = np.random.default_rng(7)
    return (rng.random((V_FULL, F_FULL, 2), dtype=np.float32) - 0.5) * 0.5


def test_full_size_silence_is_exact_zero(gas, full_src):
    with gas.Mixer(**_cfg(V_FULL, F_FULL, abi.SPEAKER_SURROUND_71)) as m:
        l, a = _setup(m, V_FULL, F_FULL, abi.SPEAKER_SURROUND_71)
        for bus in _blocks(m, V_FULL, F_FULL, np.zeros_like(full_src), 2, l, a):
            assert not bus.any()


def test_full_size_linearity_in_the_sources(gas, full_src):
    """mix(2x) == 2 mix(x): the weights do not depend on the samples (filter off), scaling by 2 is exact in fp32,
    so only the summation order (atomics) separates the two runs."""
    res = []
    for scale in (1.0, 2.0):
        with gas.Mixer(**_cfg(V_FULL, F_FULL, abi.SPEAKER_SURROUND_71)) as m:
            l, a = _setup(m, V_FULL, F_FULL, abi.SPEAKER_SURROUND_71)
            res.append(_blocks(m, V_FULL, F_FULL, full_src * np.float32(scale), 2, l, a))
    for b1, b2 in zip(*res):
        assert b1.any()
        # tolerance relative to the per-row scale: cancellation in a 16384-term sum makes per-sample relative error meaningless
        scale = np.abs(b2).max(axis=2, keepdims=True) + 1e-30
        assert np.max(np.abs(2.0 * b1.astype(np.float64) - b2) / scale) < 2e-5
        assert np.array_equal(S.routing(b1), S.routing(b2))


def test_full_size_additivity_over_voice_shards(gas, full_src):
    """The bus sum over all voices equals the sum of the buses of two disjoint halves (what the multi-GPU
    sharding relies on), each half mixed by its own context."""
    V, F, H = V_FULL, F_FULL, V_FULL // 2
    shard = gas.shard
    with gas.Mixer(**_cfg(V, F, abi.SPEAKER_SURROUND_71)) as m:
        l, a = _setup(m, V, F, abi.SPEAKER_SURROUND_71)
        whole = _blocks(m, V, F, full_src, 1, l, a)[0]
    parts = np.zeros_like(whole, dtype=np.float64)
    voices = synth.make_voices(V)
    for r in range(2):
        lo, hi = shard.instance_range(V, 2, r)
        loc, idx = shard.shard_voices(voices, V, 2, r)
        with gas.Mixer(**_cfg(H, F, abi.SPEAKER_SURROUND_71)) as m:
            inst = np.arange(H, dtype=np.int32)
            listeners = np.array([abi.identity_listener()], dtype=abi.listener)
            areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
            m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0))
            m.instance_init(inst, 0)
            em = synth.make_emitters(V, block=0, dt=F / 48000.0, area_fraction=0.25, r_min=10.0, r_max=120.0)[lo:hi].copy()
            em["instance"] -= lo
            m.gain_compute(em, listeners, areas, want_params=False)
            m.instance_start(inst)
            m.voice_init(inst)
            bus, _ = m.mix_block(loc, full_src[idx], F, want_peaks=False)
            parts += bus
    scale = np.abs(whole).max(axis=2, keepdims=True) + 1e-30
    assert np.max(np.abs(parts - whole) / scale) < 2e-5
    assert np.array_equal(S.routing(whole), S.routing(parts))


@pytest.mark.parametrize("V,F,mode", [(256, 128, abi.SPEAKER_MODE_STEREO), (300, 2048, abi.SPEAKER_SURROUND_71),
                                       (1024, 1024, abi.SPEAKER_SURROUND_31), (2048, 256, abi.SPEAKER_SURROUND_51)])
def test_sweep_corners_against_the_oracle(gas, orc, V, F, mode):
    for filt_off in (True, False):
        sc = S.default_scenario(name=f"sweep-{V}-{F}-{mode}-{filt_off}", voices=V, frames=F, speaker_mode=mode, spat=dict(mix_channel_mode=1),
                                area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, force_filter_off=filt_off, blocks=2)
        cfg = S.config_of(sc)
        with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
            got, want = S.run(m, sc, collect_state=False), S.run(o, sc, collect_state=False)
        for b, (bg, bw) in enumerate(zip(got["bus"], want["bus"])):
            assert np.array_equal(S.routing(bg), S.routing(bw))
            ok, worst, nbad = S.sample_close(bg, bw)
            assert ok, f"{sc['name']} block {b}: {nbad} samples out, worst {worst:.3e}"


def test_maximum_voice_count_runs(gas):
    """configs[4] upper corner: 65536 voices in one context (state tables, class lists, launch shapes)."""
    V, F = 65536, 128
    rng = np.random.default_rng(11)
    src = (rng.random((V, F, 2), dtype=np.float32) - 0.5) * 0.25
    with gas.Mixer(**_cfg(V, F, abi.SPEAKER_MODE_STEREO)) as m:
        l, a = _setup(m, V, F, abi.SPEAKER_MODE_STEREO)
        bus = _blocks(m, V, F, src, 1, l, a)[0]
    assert np.isfinite(bus).all() and bus.any()
