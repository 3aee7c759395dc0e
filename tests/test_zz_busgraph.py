"""The bus graph after the mix: upstream AudioServer::_mix_step's volume / mute / solo / send pass over the bus
buffers (reference README.md:98-100 steps 3-5, the demo's default_bus_layout.tres).  Restated from Godot 4.x as recalled (the engine
is not in the reference tree): pinned by known answers on the CPU; the CUDA kernel must reproduce the oracle bit for bit."""
import numpy as np
import pytest

import scenarios as S

abi, synth = S.abi, S.synth


def _db(x):
    return np.float32(np.exp(np.float32(x) * np.float32(0.11512925464970228420089957273422), dtype=np.float32))


def test_demo_layout_reverb_bus_sends_to_master(orc):
    """default_bus_layout.tres: bus 1 "Reverb", volume 0 dB, send Master."""
    rng = np.random.default_rng(0)
    bus = rng.standard_normal((2, 4, 64, 2)).astype(np.float32)
    out = orc.bus_graph(bus, [dict(), dict(volume_db=0.0, send=0)])
    np.testing.assert_array_equal(out[1], bus[1])
    np.testing.assert_array_equal(out[0], bus[0] + bus[1])


def test_volume_mute_solo_and_send_chains(orc):
    rng = np.random.default_rng(1)
    bus = rng.standard_normal((4, 2, 32, 2)).astype(np.float32)
    # 3 -> 2 -> 0, 1 -> 0; volumes applied where the signal passes
    lay = [dict(volume_db=-3.0), dict(volume_db=-6.0), dict(volume_db=2.0, send=0), dict(volume_db=-12.0, send=2)]
    out = orc.bus_graph(bus, lay)
    b3 = bus[3] * _db(-12.0)
    b2 = (bus[2] + b3) * _db(2.0)
    b1 = bus[1] * _db(-6.0)
    b0 = ((bus[0] + b2) + b1) * _db(-3.0)  # bus 2 reaches Master before bus 1 does (last bus first)
    for got, want in zip(out, (b0, b1, b2, b3)):
        np.testing.assert_array_equal(got, want.astype(np.float32))
    # mute silences the bus and everything that only reaches Master through it
    out = orc.bus_graph(bus, [dict(), dict(), dict(mute=True), dict(send=2)])
    np.testing.assert_array_equal(out[2], 0.0 * bus[2])
    np.testing.assert_array_equal(out[0], (bus[0] + out[2]) + bus[1])
    # solo: only the soloed bus and its send chain stay audible, mute flags are ignored
    out = orc.bus_graph(bus, [dict(mute=True), dict(), dict(), dict(send=2, solo=True)])
    np.testing.assert_array_equal(out[1], 0.0 * bus[1])
    np.testing.assert_array_equal(out[2], bus[2] + bus[3])
    np.testing.assert_array_equal(out[0], (bus[0] + out[2]) + out[1])
    # a send that does not point to the left goes to Master
    out = orc.bus_graph(bus, [dict(), dict(send=3), dict(send=2), dict(send=7)])
    np.testing.assert_array_equal(out[0], ((bus[0] + bus[3]) + bus[2]) + bus[1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,B", [(abi.SPEAKER_MODE_STEREO, 3), (abi.SPEAKER_SURROUND_71, 6)])
def test_cuda_bus_graph_matches_oracle_bit_for_bit(gas, orc, mode, B):
    import torch
    V, F = 128, 512
    C = mode + 1
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=B, speaker_mode=mode, mix_rate=48000.0)
    rng = np.random.default_rng(3)
    bus = rng.standard_normal((B, C, F, 2)).astype(np.float32)
    layouts = [
        [dict(volume_db=float(-2 * b), send=max(0, b - 2)) for b in range(B)],
        [dict(volume_db=1.5, mute=(b == 1), send=0 if b < 2 else b - 1) for b in range(B)],
        [dict(solo=(b == B - 1), mute=(b == 0), send=max(0, b - 1)) for b in range(B)],
    ]
    dev = torch.device("cuda", 0)
    with gas.Mixer(**cfg) as m:
        for lay in layouts:
            d_bus = torch.from_numpy(bus.copy()).to(dev)
            m.bus_layout_set(lay)
            m.bus_graph_device(d_bus.data_ptr(), F)
            m.sync()
            np.testing.assert_array_equal(d_bus.cpu().numpy(), orc.bus_graph(bus, lay))
            np.testing.assert_array_equal(m.bus_graph(bus), orc.bus_graph(bus, lay))  # host-pointer form
        with pytest.raises(gas.GasError):
            m.bus_layout_set(layouts[0][:-1])  # one descriptor per bus of the context
