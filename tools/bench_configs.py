#!/usr/bin/env python
"""Secondary BASELINE.json configurations on one B200 (not bench lines: parity-test cases that are also timed here):
  configs[1]  AudioSpatializer3D, 1024 voices, 5.1, inverse-square attenuation + attenuation filter ON (Mode B: 2C biquads/voice)
  configs[3]  AudioSpatializerEffect, 4096 voices, stereo, chain of S high-shelf stages before Master + reverb + area buses
  plus the headline workload with the filter ON (K3 instead of K2) for reference.
Prints one JSON line per configuration: us per block (CUDA-graph replay, CUDA events), voice-frames/s, per-kernel us."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402

gas, abi, synth = S.gas, S.abi, S.synth


def run(name, V, F, speaker_mode, num_buses, sc_kw, steps=200, sets=4):
    sc = S.default_scenario(voices=V, frames=F, speaker_mode=speaker_mode, num_buses=num_buses, **sc_kw)
    dev = torch.device("cuda", 0)
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(**sc["area"])], dtype=abi.area) if sc["area"] else None
    dt = F / sc["mix_rate"]
    with gas.Mixer(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=num_buses, speaker_mode=speaker_mode,
                   mix_rate=sc["mix_rate"]) as m:
        m.spatializer_set(0, S.make_spatializer(sc))
        m.instance_init(inst, 0)
        ems = [synth.make_emitters(V, block=b, dt=dt, area_fraction=sc["area_fraction"]) for b in range(sets)]
        m.gain_compute(ems[0], listeners, areas, want_params=False)
        m.instance_start(inst)
        m.voice_init(inst)
        m.listeners_set(listeners)
        if areas is not None:
            m.areas_set(areas)
        d_em = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in ems]
        d_voices = torch.from_numpy(synth.make_voices(V).view(np.uint8).copy()).to(dev)
        d_src = [(torch.rand((V, F, 2), device=dev) - 0.5) * 0.5 for _ in range(sets)]
        d_bus = torch.zeros((num_buses, speaker_mode + 1, F, 2), device=dev)

        def step(s):
            m.mix_block_device(V, d_voices.data_ptr(), d_src[s].data_ptr(), V, F, F, d_bus.data_ptr())
            m.gain_compute_device(V, d_em[(s + 1) % sets].data_ptr())

        def capture():
            gs = []
            for s in range(sets):
                m.capture_begin()
                step(s)
                gs.append(m.capture_end())
            return gs

        graphs = capture()
        stream = torch.cuda.ExternalStream(m.mix_stream, device=dev)
        for k in range(20):
            m.graph_launch(graphs[k % sets])
        m.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for k in range(steps):
            m.graph_launch(graphs[k % sets])
        with torch.cuda.stream(stream):
            e1.record()
        m.sync()
        us = 1e3 * e0.elapsed_time(e1) / steps
        m.profile_enable(True)
        pg = capture()
        for k in range(32):
            m.graph_launch(pg[k % sets])
        prof = m.profile_read()
        m.profile_enable(False)
    line = {"config": name, "voices": V, "frames": F, "channel_pairs": speaker_mode + 1, "buses": num_buses, "us_per_block": us,
            "voice_frames_per_s": V * F / (us * 1e-6), "x_realtime": (F / sc["mix_rate"]) / (us * 1e-6),
            "kernels_us": {k: 1e3 * v[0] / max(1, v[1]) for k, v in prof.items()}}
    print(json.dumps(line), flush=True)


def main():
    only = set(int(a) for a in sys.argv[1:])  # optional: indices of the configurations to run
    cfgs = [
        ("configs[1] 3D 1024 voices 5.1 inverse-square + attenuation filter (Mode B)", 1024, 512, abi.SPEAKER_SURROUND_51, 2,
         dict(spat=dict(mix_channel_mode=1, attenuation_model=abi.ATTENUATION_INVERSE_SQUARE_DISTANCE), area=dict(reverb_bus=1, amount=0.5),
              area_fraction=0.25), 200),
        ("configs[3] Effect 4096 voices stereo, 1-stage high-shelf chain, Master + reverb + area bus", 4096, 512, abi.SPEAKER_MODE_STEREO, 3,
         dict(effect_chain=[dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=1)], effect_gain_binding=0,
              area=dict(reverb_bus=2, amount=0.4, override_bus=True, bus=1), area_fraction=0.5), 200),
        ("configs[3] Effect 4096 voices stereo, 4-stage high-shelf chain, Master + reverb + area bus", 4096, 512, abi.SPEAKER_MODE_STEREO, 3,
         dict(effect_chain=[dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=4)], effect_gain_binding=0,
              area=dict(reverb_bus=2, amount=0.4, override_bus=True, bus=1), area_fraction=0.5), 200),
        ("headline workload with the attenuation filter ON (16384 voices 7.1, Mode B: 8 biquads per voice-frame)", 16384, 512,
         abi.SPEAKER_SURROUND_71, 2, dict(spat=dict(mix_channel_mode=1), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25), 50),
        ("headline workload, Mode A with the attenuation filter ON (2 biquads per voice-frame)", 16384, 512,
         abi.SPEAKER_SURROUND_71, 2, dict(spat=dict(mix_channel_mode=0), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25), 50),
    ]
    for k, (name, V, F, mode, buses, kw, steps) in enumerate(cfgs):
        if not only or k in only:
            run(name, V, F, mode, buses, kw, steps=steps)


if __name__ == "__main__":
    main()
