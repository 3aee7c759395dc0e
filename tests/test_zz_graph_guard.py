"""A captured graph has the AudioServer globals (speaker mode, mix rate, panning strength) baked into its kernel arguments: after
one of the setters it is refused instead of silently mixing with the old values."""
import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu
abi, synth = S.abi, S.synth


def test_graph_captured_before_a_global_setter_is_refused(gas):
    import torch
    V, F = 64, 128
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, num_buses=2, speaker_mode=abi.SPEAKER_SURROUND_51, mix_rate=48000.0)
    dev = torch.device("cuda", 0)
    voices = torch.from_numpy(synth.make_voices(V).view(np.uint8).copy()).to(dev)
    src = torch.zeros((V, F, 2), device=dev)
    bus = torch.zeros((2, 4, F, 2), device=dev)
    inst = np.arange(V, dtype=np.int32)
    with gas.Mixer(**cfg) as m:
        m.spatializer_set(0, abi.spatializer_defaults(mix_channel_mode=1))
        m.instance_init(inst, 0)
        m.instance_start(inst)
        m.voice_init(inst)

        def capture():
            m.capture_begin()
            m.mix_block_device(V, voices.data_ptr(), src.data_ptr(), V, F, F, bus.data_ptr())
            return m.capture_end()

        g = capture()
        m.graph_launch(g)
        m.sync()
        m.set_mix_rate(44100.0)
        with pytest.raises(gas.GasError) as ei:
            m.graph_launch(g)
        assert ei.value.status == 4  # GAS_ERR_STATE
        g2 = capture()
        m.graph_launch(g2)
        m.sync()
