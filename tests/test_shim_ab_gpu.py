"""A/B inside the stand-in engine: the SAME scene, played by the SAME reference plumbing (AudioStreamPlayerSpatial,
AudioSpatializerInstance::update_spatializer_parameters / _mix_from_playback_list, AudioServer mix step of oracle/godot_lite),
once with the reference's AudioSpatializer3D (CPU) and once with AudioSpatializer3DGPU from integration/godot_module
(gain kernel + batched mix on the B200, result fed back to the buses through feeder playbacks).

This is the closest thing to BASELINE.json configs[0] ("64 voices, stereo, 512-frame blocks at 48 kHz, headless Godot on CPU
(reference path)") that can run without an engine tree: everything module-side is the reference's own code, everything
engine-side is the godot-lite stand-in on both arms."""
import numpy as np
import pytest

import scenarios as S
from oracle import ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.gpu_available(), reason="oracle/_ref/libgas_ref_gpu.so not built")]
abi, synth = S.abi, S.synth


def _play(mixer, sc):
    # one silent AudioServer step first: the shim's feeder playbacks (registered when its context is created) fade in from
    # zero in their first step like every new playback does; a running engine has long done that when a sound starts
    V, F = sc["voices"], sc["frames"]
    mixer.spatializer_set(0, S.make_spatializer(sc))
    mixer.instance_init(np.arange(V, dtype=np.int32), 0)
    mixer.mix_block(np.zeros(0, dtype=abi.voice), np.zeros((0, F, 2), np.float32), F)
    return S.run(mixer, sc, collect_state=False)


@pytest.mark.parametrize("mode_b", [0, 1])
@pytest.mark.parametrize("speakers", [abi.SPEAKER_MODE_STEREO, abi.SPEAKER_SURROUND_51])
def test_reference_classes_vs_gpu_classes_same_scene(mode_b, speakers):
    sc = S.default_scenario(name=f"ab-{mode_b}-{speakers}", voices=64, frames=512, mix_rate=48000.0, speaker_mode=speakers,
                            spat=dict(mix_channel_mode=mode_b), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, blocks=4)
    cfg = S.config_of(sc)
    with ref.RefMixer(**cfg) as cpu:
        want = _play(cpu, sc)
    with ref.RefMixer(gpu_shim=True, **cfg) as gpu:
        got = _play(gpu, sc)
    for b, (bg, bw) in enumerate(zip(got["bus"], want["bus"])):
        assert np.abs(bw).max() > 0
        assert np.array_equal(S.routing(bg), S.routing(bw)), f"block {b}: routing differs"
        ok, worst, nbad = S.sample_close(bg, bw)
        assert ok, f"block {b}: {nbad} samples out of tolerance, worst abs err {worst:.3e}"
