#!/usr/bin/env python
"""Randomised parity campaign: seeded random scenarios (speaker mode, Mode A / B / effect chains, attenuation models, filter on /
off, areas with and without uniformity, overriding buses, two listeners, Doppler, polyphony, late starts, silent rows, peaks,
odd block sizes) played on the CUDA mixer and on the oracle, compared like tests/test_parity_gpu.py does.

    python tools/fuzz_parity.py --cases 200 --seed 1                      # on a B200
    python tests/emu/run_emulated.py tools/fuzz_parity.py --cases 200     # on the CPU emulation of the library

Prints one line per failing case (with the scenario, so it can be replayed) and a summary; exit status 1 if any case failed.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402
from oracle import orc  # noqa: E402

gas, abi = S.gas, S.abi
GAIN_RTOL = 2e-6


def random_scenario(rng, case):
    mode = int(rng.integers(0, 4))
    sc = dict(name=f"fuzz-{case}", speaker_mode=mode, seed0=int(rng.integers(0, 1000)))
    sc["voices"] = int(rng.choice([1, 3, 17, 64, 130, 257, 600]))
    sc["voices_per_instance"] = int(rng.choice([1, 1, 1, 2, 3]))
    sc["frames"] = int(rng.choice([2, 34, 64, 128, 250, 512, 514, 1024]))
    sc["blocks"] = int(rng.integers(1, 5))
    sc["num_buses"] = int(rng.choice([1, 2, 3, 5]))
    sc["mix_rate"] = float(rng.choice([44100.0, 48000.0, 22050.0]))
    spat = dict(mix_channel_mode=int(rng.integers(0, 2)), attenuation_model=int(rng.integers(0, 4)),
                unit_size=float(rng.choice([0.5, 1.0, 10.0, 40.0])), max_distance=float(rng.choice([0.0, 0.0, 30.0, 200.0])),
                panning_strength=float(rng.choice([0.0, 0.5, 1.0, 1.0, 2.0, 1.5, 0.37])),
                attenuation_filter_db=float(rng.choice([-24.0, -24.0, -6.0, 0.0, -80.0])),
                attenuation_filter_cutoff_hz=float(rng.choice([5000.0, 800.0, 15000.0])),
                emission_angle_enabled=int(rng.integers(0, 2)), emission_angle=float(rng.choice([45.0, 10.0, 90.0])),
                doppler_tracking=int(rng.choice([0, 0, 1, 2])))
    sc["spat"] = spat
    sc["listeners"] = str(rng.choice(["identity", "identity", "two", "rotated"]))
    if sc["num_buses"] >= 2 and rng.random() < 0.6:
        rb = int(rng.integers(1, sc["num_buses"]))
        area = dict(reverb_bus=rb, amount=float(rng.choice([0.0, 0.3, 1.0])), uniformity=float(rng.choice([0.0, 0.0, 0.6, 1.0])))
        if sc["num_buses"] >= 3 and rng.random() < 0.4:
            area.update(override_bus=True, bus=int(rng.integers(0, sc["num_buses"])))
        sc["area"] = area
        sc["area_fraction"] = float(rng.choice([0.25, 0.5, 1.0]))
    if rng.random() < 0.2:
        n_fx = int(rng.integers(1, 3))
        sc["effect_chain"] = [dict(mode=int(rng.integers(0, 8)),
                                   cutoff_hz=float(rng.choice([500.0, 4000.0])), resonance=float(rng.choice([0.5, 1.0])),
                                   gain=float(rng.choice([0.3, 1.5])), stages=int(rng.integers(1, 5))) for _ in range(n_fx)]
        sc["effect_gain_binding"] = int(rng.integers(-1, n_fx))
    sc["gain_every"] = int(rng.choice([1, 1, 2]))
    sc["force_filter_off"] = bool(rng.random() < 0.3)
    sc["want_peak_every"] = int(rng.choice([0, 0, 3, 1]))
    sc["silent_every"] = int(rng.choice([0, 0, 4]))
    sc["start_late"] = int(rng.choice([0, 0, 1, 2])) if sc["blocks"] > 2 else 0
    # (sources stay inside full scale: the absolute half of the tolerance, -110 dBFS, is meaningless for sums far above 1.0)
    sc["amplitude"] = float(rng.choice([1.0, 1.0, 1e-3, 0.25])) / max(1.0, (sc["voices"] / 64.0) ** 0.5)
    # (filters stay below Nyquist: above it the reference's own biquads diverge and amplify every rounding difference)
    nyq = 0.45 * sc["mix_rate"]
    spat["attenuation_filter_cutoff_hz"] = min(spat["attenuation_filter_cutoff_hz"], nyq)
    for fx in sc.get("effect_chain") or []:
        fx["cutoff_hz"] = min(fx["cutoff_hz"], nyq)
    return S.default_scenario(**sc)


def _close(g, w, rel=S.REL_TOL):
    """S.sample_close, with equal infinities accepted like equal NaNs (inf - inf is NaN)."""
    g, w = np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64)
    same_inf = np.isinf(g) & np.isinf(w) & (np.sign(g) == np.sign(w))
    return S.sample_close(np.where(same_inf, 0.0, g), np.where(same_inf, 0.0, w), rel=rel)


def check(got, want, sc):
    for b, (pg, pw) in enumerate(zip(got["params"], want["params"])):
        for f in ("update_parameters", "n_bus", "bus"):
            if not np.array_equal(pg[f], pw[f]):
                return f"block {b}: params.{f} differs"
        for f in ("mix_volumes", "bus_volumes", "pitch_scale", "linear_attenuation", "attenuation_filter_cutoff_hz"):
            if not np.allclose(pg[f], pw[f], rtol=GAIN_RTOL, atol=1e-9, equal_nan=True):
                return f"block {b}: params.{f} out of tolerance"
    for b, (bg, bw) in enumerate(zip(got["bus"], want["bus"])):
        if not np.array_equal(S.routing(bg), S.routing(bw)):
            return f"block {b}: routing differs"
        with np.errstate(invalid="ignore"):
            if np.nanmax(np.abs(np.where(np.isfinite(bw), bw, 0.0))) > 50.0:  # (stable scenarios stay below ~10: the amplitude is scaled to the voice count)
                # the reference's own recurrence diverges here (a filter driven outside its stable range): every rounding difference
                # is amplified without bound from now on, so only what came before is comparable
                return None
        ok, worst, nbad = _close(bg, bw)
        if not ok:
            return f"block {b}: {nbad} samples out of tolerance (worst {worst:.3e})"
    for b, (kg, kw) in enumerate(zip(got["peaks"], want["peaks"])):
        flagged = np.zeros(len(kw), dtype=bool)
        if sc["want_peak_every"]:
            flagged[:: sc["want_peak_every"]] = True
        ok, worst, nbad = _close(kg[flagged], kw[flagged])
        if not ok:
            return f"block {b}: peaks differ (worst {worst:.3e})"
    sg, sw = got["state"], want["state"]
    if not np.allclose(sg["prev_mix_volumes"], sw["prev_mix_volumes"], rtol=GAIN_RTOL, atol=1e-9, equal_nan=True):
        return "state: prev_mix_volumes"
    for f in ("ha1", "ha2", "hb1", "hb2"):
        ok, worst, _ = _close(sg["filter_processors"][f], sw["filter_processors"][f], rel=1e-4)
        if not ok:
            return f"state: filter history {f} (worst {worst:.3e})"
    ok, worst, _ = _close(sg["effect_history"], sw["effect_history"], rel=1e-4)
    if not ok:
        return f"state: effect history (worst {worst:.3e})"
    return None


def run_pipelined(m, sc):
    """The scenario through the pipelined form (gas_step_device): gains and plan of block k + 1 on the control warps of the launch that
    streams block k; device-resident inputs, four rotating bus / peak buffers, half of the run replayed from captured graphs.
    Returns per-block bus buffers and peaks.  (Scenarios without late starts / parameter overrides.)"""
    import torch
    V, F, vpi = sc["voices"], sc["frames"], sc["voices_per_instance"]
    n_inst = (V + vpi - 1) // vpi
    C = sc["speaker_mode"] + 1
    inst = np.arange(n_inst, dtype=np.int32)
    dev = torch.device("cuda", 0)
    listeners = S._listeners(sc)
    areas = np.array([S.synth.reverb_area(n_listeners=len(listeners), **sc["area"])], dtype=abi.area) if sc["area"] is not None else None
    dt = F / sc["mix_rate"]
    blocks = sc["blocks"]
    ems_h = [S.synth.make_emitters(n_inst, block=b, dt=dt, area_fraction=sc["area_fraction"], seed0=sc["seed0"]) for b in range(blocks)]
    voices_h = S.synth.make_voices(V, voices_per_instance=vpi)
    if sc["want_peak_every"]:
        voices_h["flags"][:: sc["want_peak_every"]] |= abi.VOICE_WANT_PEAK
    if sc["silent_every"]:
        voices_h["src_row"][(voices_h["voice"] % sc["silent_every"]) == (sc["silent_every"] - 1)] = -1
    src_h = [S.synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"], voice0=sc["seed0"]) * np.float32(sc["amplitude"]) for b in range(blocks)]
    m.spatializer_set(0, S.make_spatializer(sc))
    m.instance_init(inst, 0)
    m.gain_compute(ems_h[0], listeners, areas, want_params=False)
    m.instance_start(inst)
    m.voice_init(np.arange(V, dtype=np.int32))
    m.listeners_set(listeners)
    if areas is not None:
        m.areas_set(areas)
    d_voices = torch.from_numpy(voices_h.view(np.uint8).copy()).to(dev)
    d_ems = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in ems_h]
    d_src = [torch.from_numpy(x).to(dev) for x in src_h]
    busb = [torch.full((sc["num_buses"], C, F, 2), 7.0, device=dev) for _ in range(4)]
    peakb = [torch.full((V, 2), 7.0, device=dev) for _ in range(4)]

    def nxt(b):
        return dict(n_emitters=n_inst, d_emitters=d_ems[b].data_ptr(), n_voices=V, d_voices=d_voices.data_ptr(), src_rows=V, frames=F,
                    d_bus_out=busb[b % 4].data_ptr(), d_peaks=peakb[b % 4].data_ptr())

    out = dict(bus=[], peaks=[])
    m.step_device(next=nxt(0))
    for b in range(blocks):
        nx = nxt(b + 1) if b + 1 < blocks else None
        if b % 2 == 1 and nx is not None:
            m.capture_begin()
            m.step_device(d_src[b].data_ptr(), F, next=nx)
            g = m.capture_end()
            m.graph_launch(g)
            m.sync()
            m.graph_destroy(g)
        else:
            m.step_device(d_src[b].data_ptr(), F, next=nx)
        m.step_join_device()
        m.sync()
        out["bus"].append(busb[b % 4].cpu().numpy().copy())
        out["peaks"].append(peakb[b % 4].cpu().numpy().copy())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pipelined", action="store_true", help="play the scenarios through gas_step_device instead of the block calls")
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--first", type=int, default=0)
    args = ap.parse_args()
    bad = 0
    for case in range(args.first, args.first + args.cases):
        rng = np.random.default_rng([args.seed, case])
        sc = random_scenario(rng, case)
        cfg = S.config_of(sc)
        try:
            if args.pipelined:
                sc.update(start_late=0, force_filter_off=False, gain_every=1)
                with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
                    got = run_pipelined(m, sc)
                    want = S.run(o, sc)
                got["params"], got["state"] = want["params"], want["state"]  # (compared by the block-call campaign)
            else:
                with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
                    got = S.run(m, sc)
                    want = S.run(o, sc)
            why = check(got, want, sc)
        except Exception as ex:  # noqa: BLE001
            why = f"exception: {ex!r}"
        if why:
            bad += 1
            print(f"FAIL case {case}: {why}\n     {json.dumps(sc, default=float)}", flush=True)
    print(f"fuzz: {args.cases - bad} / {args.cases} cases match the oracle (seed {args.seed}, first {args.first})")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
