"""Scenario driver shared by the parity tests, the golden-vector generator and smoke().

A scenario is a plain dict; ``run(mixer, sc)`` plays it block by block on any object with the Mixer
surface (the CUDA Mixer or the oracle's OracleMixer) and returns everything observable: per-block
computed parameters, bus buffers, peaks, and the final voice state.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import gaspkg  # noqa: E402

gas = gaspkg.load()
abi, synth = gas.abi, gas.synth

# north-star tolerance: samples within 1e-5 relative or below -110 dBFS absolute
REL_TOL = 1e-5
ABS_TOL = 10.0 ** (-110.0 / 20.0)


def default_scenario(**kw):
    sc = dict(
        name="default",
        voices=64, voices_per_instance=1, frames=512, blocks=3, mix_rate=48000.0,
        speaker_mode=abi.SPEAKER_MODE_STEREO, num_buses=2,
        spat=dict(),                 # overrides of abi.spatializer_defaults
        listeners="identity",        # "identity" | "two" | list of abi.listener
        area=None,                   # dict(kwargs of synth.reverb_area) or None
        area_fraction=0.0,
        gain_every=1,                # recompute gains every k blocks
        force_filter_off=False,      # overwrite linear_attenuation with 0 via params_set after gain_compute
        want_peak_every=0,           # flag every k-th voice GAS_VOICE_WANT_PEAK (0 = none)
        silent_every=0,              # every k-th voice has src_row = -1 (0 = none)
        effect_chain=None,           # list of dict(mode,cutoff_hz,resonance,gain,stages) => EFFECT kind
        effect_gain_binding=-1,
        amplitude=1.0,               # extra source scale
        start_late=0,                # voices with index % 5 == 4 start at this block (0 = all at block 0)
        seed0=0,
    )
    sc.update(kw)
    return sc


def _listeners(sc):
    if isinstance(sc["listeners"], str):
        if sc["listeners"] == "identity":
            return np.array([abi.identity_listener()], dtype=abi.listener)
        if sc["listeners"] == "two":
            return np.array([abi.identity_listener(), synth.rotated_listener()], dtype=abi.listener)
        if sc["listeners"] == "rotated":
            return np.array([synth.rotated_listener()], dtype=abi.listener)
        raise ValueError(sc["listeners"])
    return np.asarray(sc["listeners"], dtype=abi.listener)


def make_spatializer(sc):
    s = abi.spatializer_defaults(**sc["spat"])
    if sc["effect_chain"] is not None:
        s["kind"] = abi.SPATIALIZER_EFFECT
        s["effect_gain_binding"] = sc["effect_gain_binding"]
        ch = np.zeros((), dtype=abi.effect_chain)
        ch["n_effects"] = len(sc["effect_chain"])
        for k, fx in enumerate(sc["effect_chain"]):
            ch["effects"][k]["mode"] = fx.get("mode", abi.FILTER_HIGHSHELF)
            ch["effects"][k]["cutoff_hz"] = fx.get("cutoff_hz", 5000.0)
            ch["effects"][k]["resonance"] = fx.get("resonance", 1.0)
            ch["effects"][k]["gain"] = fx.get("gain", 0.25)
            ch["effects"][k]["stages"] = fx.get("stages", 1)
        s["chain"] = ch
    return s


def config_of(sc):
    n_inst = (sc["voices"] + sc["voices_per_instance"] - 1) // sc["voices_per_instance"]
    return dict(max_instances=max(n_inst, 1), max_voices=max(sc["voices"], 1), max_frames=sc["frames"], max_spatializers=2,
                num_buses=sc["num_buses"], speaker_mode=sc["speaker_mode"], mix_rate=sc["mix_rate"])


def run(mixer, sc, collect_state=True):
    V, F = sc["voices"], sc["frames"]
    vpi = sc["voices_per_instance"]
    n_inst = (V + vpi - 1) // vpi
    inst = np.arange(n_inst, dtype=np.int32)
    mixer.spatializer_set(0, make_spatializer(sc))
    mixer.instance_init(inst, 0)
    listeners = _listeners(sc)
    areas = None
    if sc["area"] is not None:
        areas = np.array([synth.reverb_area(n_listeners=len(listeners), **sc["area"])], dtype=abi.area)
    dt = F / sc["mix_rate"]
    voices_all = synth.make_voices(V, voices_per_instance=vpi)
    if sc["want_peak_every"]:
        voices_all["flags"][:: sc["want_peak_every"]] |= abi.VOICE_WANT_PEAK
    late = np.zeros(V, dtype=bool)
    if sc["start_late"]:
        late[4::5] = True
    started = np.zeros(V, dtype=bool)
    out = dict(params=[], bus=[], peaks=[])
    for b in range(sc["blocks"]):
        if b % sc["gain_every"] == 0:
            em = synth.make_emitters(n_inst, block=b, dt=dt, area_fraction=sc["area_fraction"], seed0=sc["seed0"])
            p = mixer.gain_compute(em, listeners, areas)
            if sc["force_filter_off"]:
                p = p.copy()
                p["linear_attenuation"] = 0.0
                mixer.params_set(inst, p)
            out["params"].append(p)
        if b == 0:
            mixer.instance_start(inst)  # after the first parameters exist, like the physics tick does
        live = ~late if b < sc["start_late"] else np.ones(V, dtype=bool)
        fresh = live & ~started
        if fresh.any():
            mixer.voice_init(np.nonzero(fresh)[0].astype(np.int32))
            started |= fresh
        src = synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"], voice0=sc["seed0"]) * np.float32(sc["amplitude"])
        voices = voices_all[live].copy()
        if sc["silent_every"]:
            sil = (voices["voice"] % sc["silent_every"]) == (sc["silent_every"] - 1)
            voices["src_row"][sil] = -1
        bus, peaks = mixer.mix_block(voices, src, F)
        out["bus"].append(bus)
        full = np.zeros((V, 2), dtype=np.float32)
        full[live] = peaks
        out["peaks"].append(full)
    if collect_state:
        out["state"] = mixer.voice_state_export(np.arange(V, dtype=np.int32))
    return out


def sample_close(got, want, rel=REL_TOL, abs_tol=ABS_TOL):
    """north-star sample criterion: |g-w| <= rel*|w|  or  |g-w| < abs_tol.  Returns (ok, worst, n_bad).
    A NaN passes only against a NaN in the same place (the reference produces NaN gains itself)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    both_nan = np.isnan(got) & np.isnan(want)
    with np.errstate(invalid="ignore"):
        err = np.abs(got - want)
        ok = (err <= rel * np.abs(want)) | (err < abs_tol) | both_nan
    finite = err[np.isfinite(err)]
    worst = float(finite.max()) if finite.size else 0.0
    return bool(ok.all()), worst, int((~ok).sum())


def routing(bus):
    """Boolean [bus, pair, side] pattern of non-silent outputs: the bit-exact routing gate (a NaN counts as signal)."""
    return (np.nan_to_num(np.abs(np.asarray(bus)), nan=1.0).max(axis=2) > 0)
