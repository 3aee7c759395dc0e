#!/usr/bin/env python
"""Generates tests/golden/*.npz — committed golden input/output vectors for the hot path.

Provenance: the vectors are OUTPUTS OF THE REFERENCE ITSELF, run in this container: oracle/_ref is
/root/reference/*.cpp, unmodified, compiled against the godot-lite stand-in headers (oracle/godot_lite/,
oracle/Makefile target `ref`) and driven by oracle/ref_harness.cpp.  Parameters, bus buffers and the final
previous-volume state come from that library; `peaks` come from the CPU oracle (the reference keeps a voice's block
peak in a local variable, audio_spatializer.cpp:419 — it is observable only through deactivation, which
tests/test_lifecycle.py covers).  The generator refuses to write a file unless the CPU oracle reproduces the
reference's numbers bit for bit.  The reference cannot travel to the GPU box, these files can.
Inputs are not stored: they are a pure function of the scenario dict (splitmix64 generator in
godot-audio-spatializer_b200/synth.py), which is stored.

Usage:  python tests/golden/make_golden.py        (rewrites the .npz files; needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import scenarios as S  # noqa: E402

abi = S.abi

# BASELINE.json configs, scaled to sizes the oracle finishes in well under a second
SCENARIOS = {
    # configs[0]: 64 voices, stereo, 512-frame blocks @ 48 kHz, both modes
    "cfg0_stereo_mode_a": dict(voices=64, frames=512, speaker_mode=abi.SPEAKER_MODE_STEREO, spat=dict(mix_channel_mode=0), blocks=3),
    "cfg0_stereo_mode_b": dict(voices=64, frames=512, speaker_mode=abi.SPEAKER_MODE_STEREO, spat=dict(mix_channel_mode=1), blocks=3),
    # configs[1]: 5.1, inverse-square attenuation + attenuation filter, Mode B (voice count scaled 1024 -> 96)
    "cfg1_51_invsq_filter": dict(voices=96, frames=512, speaker_mode=abi.SPEAKER_SURROUND_51,
                                 spat=dict(mix_channel_mode=1, attenuation_model=abi.ATTENUATION_INVERSE_SQUARE_DISTANCE), blocks=3,
                                 want_peak_every=8),
    # configs[2]: 7.1, filter off, Master + reverb bus (voice count scaled 16384 -> 128)
    "cfg2_71_stream_reverb": dict(voices=128, frames=512, speaker_mode=abi.SPEAKER_SURROUND_71,
                                  spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0),
                                  area=dict(reverb_bus=1, amount=0.5, uniformity=0.0), area_fraction=0.25, blocks=3),
    # configs[3]: AudioSpatializerEffect, biquad chain before multi-bus sends (voice count scaled 4096 -> 64)
    "cfg3_effect_chain": dict(voices=64, frames=512, speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=3,
                              effect_chain=[dict(mode=abi.FILTER_HIGHSHELF, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=2),
                                            dict(mode=abi.FILTER_LOWPASS, cutoff_hz=9000.0, resonance=0.7, gain=1.0, stages=1)],
                              effect_gain_binding=0, area=dict(reverb_bus=2, amount=0.4, uniformity=0.5), area_fraction=0.5, blocks=3),
    # edge cases: two listeners, max distance, emission angle, polyphony, late starts, silent tails, odd block size
    "edge_polyphony_odd": dict(voices=45, voices_per_instance=3, frames=130, speaker_mode=abi.SPEAKER_SURROUND_31,
                               spat=dict(mix_channel_mode=1, max_distance=100.0, emission_angle_enabled=1, emission_angle=30.0),
                               listeners="two", blocks=4, start_late=2, silent_every=7, want_peak_every=5),
}


def scenario(name):
    return S.default_scenario(name=name, **SCENARIOS[name])


def run_oracle(name):
    from oracle import orc
    sc = scenario(name)
    with orc.OracleMixer(**S.config_of(sc)) as o:
        return S.run(o, sc)


def run_reference(name):
    """The same scenario through the reference module's own code (oracle/_ref); peaks filled in from the oracle."""
    from oracle import ref
    sc = scenario(name)
    with ref.RefMixer(**S.config_of(sc)) as r:
        out = S.run(r, sc)
    out["peaks"] = run_oracle(name)["peaks"]
    return out


def pack(out):
    d = {"bus": np.stack(out["bus"]), "peaks": np.stack(out["peaks"])}
    p = np.stack(out["params"])
    for f in ("mix_volumes", "bus_volumes", "pitch_scale", "linear_attenuation", "attenuation_filter_cutoff_hz", "update_parameters",
              "n_bus", "bus"):
        d["params_" + f] = np.ascontiguousarray(p[f])
    d["state_prev_mix_volumes"] = np.ascontiguousarray(out["state"]["prev_mix_volumes"])
    return d


def main():
    for name in SCENARIOS:
        d = pack(run_reference(name))
        o = pack(run_oracle(name))
        for k in d:
            if not np.array_equal(d[k], o[k], equal_nan=True):
                raise SystemExit(f"{name}: the oracle does not reproduce the reference's {k}; not writing")
        d["scenario_json"] = np.array(json.dumps({k: (v if not isinstance(v, np.generic) else v.item()) for k, v in SCENARIOS[name].items()},
                                                 default=lambda o: o.item() if hasattr(o, "item") else str(o)))
        d["minted_by"] = np.array("oracle/_ref: /root/reference/*.cpp (unmodified) + oracle/godot_lite + oracle/ref_harness.cpp")
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: bus {d['bus'].shape}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
