#!/usr/bin/env python
"""bench.py — mixed voice-frames/sec of the batched spatial mixer on N B200s (one process per GPU).

One *step* = one pass of the hot path over one mix block of synthetic input on every rank: gains (calculate_spatialization for
every instance), plan (what process_frames / mix_channel and AudioServer decide per voice before their sample loops), ramped
mix of every voice into the bus buffers, and (N > 1) the sum of the per-GPU partial bus buffers.  A step is ONE launch of the step
kernel (gas_step_device: it streams block k while its control warps compute gains and plan of block k + 1).  At 4 and 8 GPUs — where
that form never completed a run on hardware this round — it is measured in a watchdog-guarded child process per rank
(guarded_pipelined_attempt); if a child stops making progress all are killed and the ranks measure with the block-call form
(gas_mix_block_device + gas_gain_compute_device) instead, see select_form().  Workload (BASELINE.json
configs[2] shape): AudioSpatializer3D, 16384 voices per GPU x 512-frame blocks at 48 kHz, 7.1 (4 channel pairs),
mix_channel_mode on, two buses (Master + a reverb bus fed by the voices inside a reverb Area3D), attenuation filter inactive
because the reference skips it below 0.001 linear gain (audio_spatializer_3d.cpp:568) — obtained with attenuation_filter_db = -80
and unit_size = 1, no override.

value      whole-job voice-frames/s, inputs resident in HBM, CUDA-graph replay of the device-resident C-ABI calls, timed with CUDA
           events on the mix stream, max over ranks.
e2e        same metric through the public API with HOST inputs.  With device-resident sources (gas_mix_block_resident: PCM clips in
           HBM, resampler + voice lifecycle + mix on the device) only the emitters and the voice list go up per step and the bus
           buffers come back; that leg runs in a child process and is adopted only if its own parity check against the oracle
           passes.  Otherwise (and always as e2e.host_frames) gas_gain_compute + gas_mix_block with pinned host source frames:
           68 MB up per step, PCIe-bound.
roofline   the step kernel: algorithmic bytes per launch / its average launch duration over the timed region (launches overlap, so
           that is the step time) vs the measured HBM copy peak in MEASURED_PEAKS.json; the isolated reading between event-record
           nodes beside it.
cpu_baseline  the CPU oracle (restated reference loop, oracle/, pinned by the reference's own sources in oracle/_ref) on the host
           cores, bounded sample.

configs    (N = 1) the other BASELINE.json configurations, time-boxed: us per block, voice-frames/s, fraction of the HBM
           and fp32 rooflines, CPU baseline and a full-size parity flag each.

`--impl reference` times the reference's CPU implementation of the path on all host threads: the oracle port (the
reference module itself is compiled and executed only as the checker that pins the oracle, oracle/_ref — it is a
single-threaded object graph, not a batch mixer; headless Godot cannot be built here).
"""
import hashlib
import argparse
import math
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "mixed voice-frames/sec"
UNIT = "voice-frames/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

WORKLOAD = dict(voices=16384, frames=512, mix_rate=48000.0, speaker_mode=3, num_buses=2, area_fraction=0.25,
                spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0), r_min=10.0, r_max=120.0)
N_SETS = 8  # distinct source / emitter sets rotated through: 8 x 64 MiB = 512 MiB >= 4 x L2
NB = 4      # bus buffers in rotation
CLASSIC = bool(os.environ.get("GAS_BENCH_CLASSIC"))  # experiments: one gas_mix_block_device + gas_gain_compute_device call per step


def workload_name(w, filt="off"):
    return (f"AudioSpatializer3D {w['voices']} voices/GPU x {w['frames']}-frame blocks @ {int(w['mix_rate'])} Hz, 7.1, "
            f"mix_channel_mode, Master + reverb bus ({int(w['area_fraction'] * 100)}% of voices in the reverb area), filter {filt}")


def algorithmic_bytes(V, F, C, B_out, filter_on=False):
    """8 V F (source read) + V S(C) (per-voice parameters/state) + 8 F C B_out (bus write)."""
    S = 24 + (168 if filter_on else 24) * C
    return 8 * V * F + V * S + 8 * F * C * B_out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=None):
        super().__init__(daemon=True)
        # NVML queries are not free for the GPU they ask about: a few samples inside the timed region, not hundreds
        self.index, self.period = index, float(os.environ.get("GAS_BENCH_SAMPLE_PERIOD", "0.02")) if period is None else period
        self.samples, self._stop_evt = [], threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.time(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        seen = set()
        for _, _, r in inside:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        mhz = sorted(s[1] for s in inside)
        return {"sm_mhz": (mhz[len(mhz) // 2] if mhz else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(seen),
                "samples": len(inside)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ---------------------------------------------------------------------------------------------------------------
def build_host_inputs(w, abi, synth, n_emitter_sets=2, src_sets=1):
    V, F = w["voices"], w["frames"]
    dt = F / w["mix_rate"]
    emitters = [synth.make_emitters(V, block=b, dt=dt, area_fraction=w["area_fraction"], r_min=w["r_min"], r_max=w["r_max"])
                for b in range(n_emitter_sets)]
    srcs = [synth.make_sources(V, F, block=b, mix_rate=w["mix_rate"]) for b in range(src_sets)]
    voices = synth.make_voices(V)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
    return emitters, srcs, voices, listeners, areas


def setup_mixer(m, w, abi, emitters, listeners, areas):
    V = w["voices"]
    inst = np.arange(V, dtype=np.int32)
    m.spatializer_set(0, abi.spatializer_defaults(**w["spat"]))
    m.instance_init(inst, 0)
    m.gain_compute(emitters[0], listeners, areas, want_params=False)
    m.instance_start(inst)
    m.voice_init(inst)


def run_cpu(w, steps, warmup, threads, abi, synth, budget_s=None, inputs=None):
    """Oracle port of the reference loop: gain (serial, like the physics thread) + mix per step.
    Returns (voice-frames/s, steps actually timed, seconds)."""
    from oracle import orc
    V, F = w["voices"], w["frames"]
    emitters, srcs, voices, listeners, areas = inputs or build_host_inputs(w, abi, synth)
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=w["num_buses"], speaker_mode=w["speaker_mode"],
               mix_rate=w["mix_rate"])
    with orc.OracleMixer(**cfg) as o:
        setup_mixer(o, w, abi, emitters, listeners, areas)
        for k in range(warmup):
            o.gain_compute(emitters[k % len(emitters)], listeners, areas, want_params=False)
            o.mix_block(voices, srcs[k % len(srcs)], F, want_peaks=False, threads=threads)
        t0 = time.perf_counter()
        done = 0
        for k in range(steps):
            o.gain_compute(emitters[k % len(emitters)], listeners, areas, want_params=False)
            o.mix_block(voices, srcs[k % len(srcs)], F, want_peaks=False, threads=threads)
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return V * F * done / dt, done, dt


def host_threads():
    """Host threads the CPU arms may use: the affinity mask, not OMP_NUM_THREADS (torchrun forces that to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bench_config(w, world, peer=True):
    """The `config` object of the JSON line: identical in the GPU arm and the reference arm."""
    V, F, C, B = w["voices"], w["frames"], w["speaker_mode"] + 1, w["num_buses"]
    return {"workload": workload_name(w), "voices_per_gpu": V, "frames": F, "channel_pairs": C, "buses": B,
            "l2": f"{N_SETS} distinct source sets of {V * F * 8 / 2**20:.0f} MiB rotated (> 4x L2)",
            "launch": ("CUDA-graph replay of gas_step_device: one kernel launch per step streams block k and computes gains + plan of "
                       "block k+1 on its control warps; up to 8 consecutive steps per graph launch" if not CLASSIC else
                       "CUDA-graph replay of gas_mix_block_device (block k) with gas_gain_compute_device (parameters of block "
                       "k+1) beside it on the gain stream; up to 8 consecutive steps per graph launch"),
            "reduce": ("none (1 GPU)" if world == 1 else
                       "gas_reduce_bus_exchange_device inside the step graph, one block in flight on the exchange stream: every rank "
                       "adds its partial bus buffer into every rank's exchange buffer with vector reductions on peer pointers "
                       "(NVLink), one arrival-counter round per block"
                       if peer else "torch.distributed all_reduce (NCCL) of the partial bus buffers, one call per step")}


def select_form(world):
    """Pipelined form (gas_step_device) at 1 and 2 GPUs.  At 4 and 8 GPUs main() first tries it in guarded child processes
    (GAS_BENCH_PIPELINED set for them); the process that lands here without that variable is the fallback and uses the block-call
    form, whose kernels never wait for each other's CTAs (the structure that ran on 8 GPUs in round 1).
    Background: the pipelined form was validated on hardware at 1 and 2 GPUs; the only 4- / 8-GPU run of the round hung (an NCCL
    barrier enqueued while step kernels were in flight, see barrier() in gpu_arm) and took the rest of the round's GPU budget with
    it, so the fix has only run on the CPU emulation of the library (tests/emu, 4 and 8 emulated ranks) since."""
    global CLASSIC
    if world >= 4 and not os.environ.get("GAS_BENCH_PIPELINED"):
        CLASSIC = True


def reference_arm(args):
    rank, _, world = dist_env()
    select_form(max(world, args.gpus))
    if rank != 0:
        return
    import gaspkg
    gas = gaspkg.load()
    abi, synth = gas.abi, gas.synth
    w = dict(WORKLOAD)
    threads = host_threads()
    # every step is a bounded sample of the workload: one full block of one GPU's share (16384 voices) on the CPU
    steps = max(1, args.steps)
    warmup = max(1, args.warmup)
    value, done, secs = run_cpu(w, steps, warmup, threads, abi, synth, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": warmup,
        "ms_per_step": 1e3 * secs / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(w, max(1, args.gpus), args.reduce == "peer"),
        "note": "reference CPU path = oracle port (gcc -O2, OpenMP over instances), pinned bit for bit against the reference module's "
                "own code by oracle/_ref; headless Godot cannot be built here (needs the engine tree + scons)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{done} full blocks of {w['voices']} voices x {w['frames']} frames (gain + mix)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# 4 / 8 GPUs: the pipelined form under a watchdog
# ---------------------------------------------------------------------------------------------------------------
def heartbeat(what=""):
    """Child of guarded_pipelined_attempt: tells the parent that the main thread is alive (one file per rank)."""
    path = os.environ.get("GAS_BENCH_HB_FILE")
    if path:
        try:
            with open(path, "w") as f:
                f.write(f"{time.time():.3f} {what}\n")
        except OSError:
            pass
    if path and os.environ.get("GAS_BENCH_TEST_HANG") == what and dist_env()[0] == 1:  # tests: rank 1's child stops here for good
        time.sleep(3600)


def guarded_pipelined_attempt(args, hb_timeout=None, total_timeout=None):
    """Runs this very command once more in a child process per rank with the pipelined form forced (own rendezvous port), under a
    watchdog: a rank whose child stops reporting progress kills it and says so, every other rank then kills its own.  True = all
    children finished and rank 0 has printed the child's line; False = the caller measures with the block-call form instead.

    Why: the pipelined form was validated on hardware at 1 and 2 GPUs only; its single 4- / 8-GPU run of the round hung (an NCCL
    barrier enqueued behind in-flight step kernels, since fixed in barrier()) and the fix could not be re-run on hardware.  A hang
    must cost the scaling run some time, never its data points."""
    import subprocess
    rank, _, world = dist_env()
    hb_timeout = hb_timeout or float(os.environ.get("GAS_BENCH_HB_TIMEOUT", "120"))
    total_timeout = total_timeout or float(os.environ.get("GAS_BENCH_ATTEMPT_TIMEOUT", "420"))
    port = int(os.environ.get("MASTER_PORT", "29500"))
    try:  # one directory per launch: the launcher's pid and start time (the ranks of one launch share both)
        with open(f"/proc/{os.getppid()}/stat") as f:
            born = f.read().rsplit(")", 1)[1].split()[19]
    except Exception:
        born = "0"
    run_dir = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"gas_bench_{os.getppid()}_{born}_{port}")
    os.makedirs(run_dir, exist_ok=True)
    hb_file = os.path.join(run_dir, f"hb_{rank}")
    out_file = os.path.join(run_dir, f"out_{rank}")
    err_file = os.path.join(run_dir, f"err_{rank}")

    def flag(r):
        return os.path.join(run_dir, f"done_{r}")

    env = dict(os.environ, MASTER_PORT=str(port + 53), GAS_BENCH_PIPELINED="1", GAS_BENCH_CHILD="1", GAS_BENCH_HB_FILE=hb_file)
    env.pop("TORCHELASTIC_USE_AGENT_STORE", None)  # the children rendezvous on a store of their own (rank 0's child hosts it)
    t0 = time.time()
    with open(hb_file, "w") as f:
        f.write(f"{t0:.3f} spawn\n")
    ok, why = False, ""
    with open(out_file, "w") as fo, open(err_file, "w") as fe:
        child = subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, stdout=fo, stderr=fe)
        while True:
            rc = child.poll()
            if rc is not None:
                ok, why = rc == 0, f"exit {rc}"
                break
            now = time.time()
            try:
                last = os.path.getmtime(hb_file)
                started = not open(hb_file).read().rstrip().endswith("spawn")
            except OSError:
                last, started = t0, False
            limit = hb_timeout if started else 2 * hb_timeout  # the first import of torch on a fresh box can take a minute
            others_failed = any(os.path.exists(flag(r)) and open(flag(r)).read().startswith("fail") for r in range(world) if r != rank)
            if now - last > limit or now - t0 > total_timeout or others_failed:
                why = "another rank gave up" if others_failed else f"no progress for {now - last:.0f} s (total {now - t0:.0f} s)"
                child.kill()
                try:
                    child.wait(timeout=30)
                except Exception:
                    pass
                break
            time.sleep(0.5)
    try:  # what the child said on stderr (NCCL's own log lines among it) belongs to this run's stderr
        with open(err_file) as f:
            sys.stderr.write(f.read()[-200000:])
        sys.stderr.flush()
    except OSError:
        pass
    with open(flag(rank) + ".tmp", "w") as f:
        f.write(("ok " if ok else "fail ") + why + "\n")
    os.replace(flag(rank) + ".tmp", flag(rank))
    # every rank decides the same way: wait for all verdicts
    t1 = time.time()
    verdicts = {}
    while len(verdicts) < world and time.time() - t1 < hb_timeout + 60:
        for r in range(world):
            if r not in verdicts and os.path.exists(flag(r)):
                verdicts[r] = open(flag(r)).read().strip()
        time.sleep(0.2)
    all_ok = len(verdicts) == world and all(v.startswith("ok") for v in verdicts.values())
    line = None
    if rank == 0:
        if all_ok:
            try:
                for ln in reversed(open(out_file).read().strip().splitlines()):
                    if ln.startswith("{"):
                        line = json.loads(ln)
                        break
            except Exception:
                line = None
            all_ok = line is not None
            if not all_ok:  # (the other ranks cannot see this: they leave, rank 0 reports the failure instead of a number)
                print(json.dumps({"error": "pipelined attempt finished without a result line", "n_gpus": world}), flush=True)
                return True
        if all_ok:
            line["form"] = "pipelined (gas_step_device), measured in a watchdog-guarded child process per rank"
            print(json.dumps(line), flush=True)
        else:
            print(f"[bench] pipelined attempt at {world} GPUs abandoned ({verdicts}); measuring with the block-call form", file=sys.stderr, flush=True)
    return all_ok


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class DeviceWorkload:
    """Everything one rank needs resident in HBM."""

    def __init__(self, gas, torch, w, device, rank, parity_src=None):
        abi, synth = gas.abi, gas.synth
        self.gas, self.torch, self.w = gas, torch, w
        V, F = w["voices"], w["frames"]
        self.C = w["speaker_mode"] + 1
        dt = F / w["mix_rate"]
        self.listeners = np.array([abi.identity_listener()], dtype=abi.listener)
        self.areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
        self.emitters_host = [synth.make_emitters(V, block=b, dt=dt, area_fraction=w["area_fraction"], r_min=w["r_min"], r_max=w["r_max"],
                                                  seed0=rank * 1000003) for b in range(N_SETS)]
        self.voices_host = synth.make_voices(V)
        self.mixer = gas.Mixer(device=device, max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=w["num_buses"],
                               speaker_mode=w["speaker_mode"], mix_rate=w["mix_rate"])
        dev = torch.device("cuda", device)
        self.d_emitters = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in self.emitters_host]
        self.d_voices = torch.from_numpy(self.voices_host.view(np.uint8).copy()).to(dev)
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        self.d_src = []
        for k in range(N_SETS):
            if k == 0 and parity_src is not None:
                self.d_src.append(torch.from_numpy(parity_src).to(dev))
            else:
                self.d_src.append((torch.rand((V, F, 2), generator=g, device=dev, dtype=torch.float32) - 0.5) * 0.5)
        # bus buffers rotate over 4: the step kernel of block k zeroes the buffer of block k + 1 while block k - 1's may still be
        # read by the exchange (N > 1) or added to by its voice-parallel kernel
        self.d_bus = [torch.zeros((w["num_buses"], self.C, F, 2), device=dev, dtype=torch.float32) for _ in range(NB)]
        self.d_sum = [torch.zeros((w["num_buses"], self.C, F, 2), device=dev, dtype=torch.float32) for _ in range(2)]  # N > 1: reduced
        self.d_gate = torch.zeros((w["num_buses"], self.C, F, 2), device=dev, dtype=torch.float32)  # N > 1: start gate scratch
        setup_mixer(self.mixer, w, abi, self.emitters_host, self.listeners, self.areas)
        self.mixer.listeners_set(self.listeners)
        self.mixer.areas_set(self.areas)
        self.mixer.sync()

    def comm_setup(self, dist):
        """Peer-memory reduce: every rank exports its exchange buffer (CUDA IPC handle), all ranks open all."""
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, self.mixer.comm_export())
        self.mixer.comm_open(dist.get_rank(), handles)
        dist.barrier()
        self.peer_reduce = True

    def reduce_prime(self):
        pass  # the first exchange finds no outstanding block and only pushes

    def reduce_drain(self, k_last):
        """Pushes the last block and ends the two blocks still in flight."""
        if getattr(self, "peer_reduce", False):
            F = self.w["frames"]
            self.mixer.step_join_device()
            self.mixer.reduce_bus_exchange_device(self.d_bus[k_last % NB].data_ptr(), self.d_sum[(k_last + 1) % 2].data_ptr(), F)
            self.mixer.reduce_bus_end_device(self.d_sum[k_last % 2].data_ptr(), F)

    def next_block(self, k):
        """What the step kernel of block k - 1 prepares: gains (emitter set of block k) and plan of block k."""
        w = self.w
        nx = dict(n_voices=w["voices"], d_voices=self.d_voices.data_ptr(), src_rows=w["voices"], frames=w["frames"],
                  d_bus_out=self.d_bus[k % NB].data_ptr())
        if not (os.environ.get("GAS_BENCH_NOGAIN") or getattr(self, "no_gain", False)):  # gains left out: experiments
            nx.update(n_emitters=w["voices"], d_emitters=self.d_emitters[k % N_SETS].data_ptr())
        return nx

    def restart(self):
        """Ends a pipelined run (streams the block still planned on the device) and plans block 0 again, so that step 0 can follow."""
        if CLASSIC:
            return
        m, w = self.mixer, self.w
        m.step_device(self.d_src[0].data_ptr(), w["frames"], next=None)
        m.step_device(next=self.next_block(0))

    def step_device(self, k):
        """One step = ONE launch of the step kernel: it streams block k into its bus buffers while its control warps compute,
        for block k + 1, the gains (calculate_spatialization of every instance: the reference's physics thread) and the plan
        (what process_frames / mix_channel and AudioServer decide per voice before their sample loops)."""
        m, w = self.mixer, self.w
        s = k % N_SETS
        if CLASSIC:
            m.mix_block_device(w["voices"], self.d_voices.data_ptr(), self.d_src[s].data_ptr(), w["voices"], w["frames"], w["frames"],
                               self.d_bus[k % NB].data_ptr())
        else:
            m.step_device(self.d_src[s].data_ptr(), w["frames"], next=self.next_block(k + 1))
        if getattr(self, "peer_reduce", False):
            # N > 1: sum of the per-GPU partial bus buffers over peer memory, inside the graph, one block in flight: while
            # block k mixes, the exchange stream pushes block k-1's partial sums to every rank and writes block k-2's
            # complete sum — the whole exchange (NVLink round trip, system-scope fences, rank skew) is off the critical path.
            m.reduce_bus_exchange_device(self.d_bus[(k - 1) % NB].data_ptr(), self.d_sum[k % 2].data_ptr(), w["frames"])
        if CLASSIC and not (os.environ.get("GAS_BENCH_NOGAIN") or getattr(self, "no_gain", False)):
            m.gain_compute_device(w["voices"], self.d_emitters[(s + 1) % N_SETS].data_ptr())
        return self.d_bus[k % NB]

    def capture_steps(self, chunk=1):
        """CUDA graphs of `chunk` consecutive steps each: graph g replays steps g*chunk .. g*chunk+chunk-1, and as many graphs are
        captured as it takes to come back to step 0 of the rotation (sources and emitters rotate over N_SETS, bus buffers over NB).
        Several steps per graph take the graph-to-graph hand-over off all but one step in `chunk`."""
        graphs = []
        self.restart()
        for g in range(graphs_per_cycle(chunk)):
            self.mixer.capture_begin()
            for j in range(chunk):
                self.step_device(g * chunk + j)
            graphs.append(self.mixer.capture_end())
        return graphs


ROTATION = N_SETS  # steps after which sources (N_SETS), emitters (N_SETS), bus buffers (NB) and reduce buffers (2) all repeat
assert ROTATION % NB == 0 and ROTATION % 2 == 0


def graphs_per_cycle(chunk):
    """Graphs of `chunk` steps needed before the rotation is back at step 0."""
    return ROTATION // math.gcd(chunk, ROTATION)


def steps_per_graph(K):
    """Largest of 8, 4, 2, 1 that divides the timed step count (GAS_BENCH_CHUNK overrides with any divisor up to 64: longer graphs
    are supported by capture_steps but were not measured — a long graph's launch latency is exposed when it is the first of the
    timed region)."""
    forced = os.environ.get("GAS_BENCH_CHUNK")
    if forced:
        c = int(forced)
        if 1 <= c <= 64 and K % c == 0:
            return c
    for c in (8, 4, 2):
        if K % c == 0:
            return c
    return 1


def parity_gate(gas, w, device, abi, synth, parity_src, host_inputs):
    """Full-size parity of the exact bench workload: 2 state-carrying blocks on source set 0 against
    the float32 oracle (north-star tolerance) and the float64 shadow (accumulation-order budget)."""
    from oracle import orc
    import scenarios as S
    V, F = w["voices"], w["frames"]
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=w["num_buses"], speaker_mode=w["speaker_mode"],
               mix_rate=w["mix_rate"])
    emitters, _, voices, listeners, areas = host_inputs
    res = {}
    with orc.OracleMixer(**cfg) as o, gas.Mixer(device=device, **cfg) as m:
        for mm in (o, m):
            setup_mixer(mm, w, abi, emitters, listeners, areas)
        for b in range(2):
            for mm in (o, m):
                mm.gain_compute(emitters[b % len(emitters)], listeners, areas, want_params=False)
            want, _ = o.mix_block(voices, parity_src, F, want_peaks=False, shadow=True, threads=host_threads())
            got, _ = m.mix_block(voices, parity_src, F, want_peaks=False)
            shadow = o.last_bus64
            ok, worst, nbad = S.sample_close(got, want)
            scale = float(np.abs(shadow).max())
            res[f"block{b}"] = {
                "routing_exact": bool(np.array_equal(S.routing(got), S.routing(want))),
                "within_1e-5_rel_or_-110dBFS_of_f32_oracle": ok, "samples_out": nbad, "worst_abs_err_vs_f32_oracle": worst,
                "max_abs_err_gpu_vs_f64_shadow": float(np.abs(got - shadow).max()),
                "max_abs_err_f32_oracle_vs_f64_shadow": float(np.abs(want - shadow).max()),
                "peak_abs_bus_sample": scale,
            }
    return res


def parity_gate_multi(gas, torch, dist, w, local_rank, rank, world, abi, synth):
    """N > 1: one block of the bench workload on every rank (its own shard of N x V voices), summed by
    gas_reduce_bus_device over peer memory; rank 0 compares the reduced bus buffers with the oracle's mix of ALL voices."""
    import scenarios as S
    V, F = w["voices"], w["frames"]
    dt = F / w["mix_rate"]
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=w["num_buses"], speaker_mode=w["speaker_mode"],
               mix_rate=w["mix_rate"])
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)

    def rank_inputs(r):
        em = synth.make_emitters(V, block=0, dt=dt, area_fraction=w["area_fraction"], r_min=w["r_min"], r_max=w["r_max"], seed0=r * 1000003)
        return em, synth.make_sources(V, F, block=0, mix_rate=w["mix_rate"], voice0=r * V)

    dev = torch.device("cuda", local_rank)
    em, src = rank_inputs(rank)
    voices = synth.make_voices(V)
    with gas.Mixer(device=local_rank, **cfg) as m:
        handles = [None] * world
        dist.all_gather_object(handles, m.comm_export())
        m.comm_open(rank, handles)
        dist.barrier()
        setup_mixer(m, w, abi, [em], listeners, areas)
        d_voices = torch.from_numpy(voices.view(np.uint8).copy()).to(dev)
        d_src = torch.from_numpy(src).to(dev)
        d_bus = torch.zeros((w["num_buses"], w["speaker_mode"] + 1, F, 2), device=dev, dtype=torch.float32)
        m.mix_block_device(V, d_voices.data_ptr(), d_src.data_ptr(), V, F, F, d_bus.data_ptr())
        m.reduce_bus_device(d_bus.data_ptr(), F)
        m.sync()
        got = d_bus.cpu().numpy()
        dist.barrier()
        m.comm_close()
    if rank != 0:
        return None
    from oracle import orc
    cfg_all = dict(cfg, max_instances=V * world, max_voices=V * world)
    ems, srcs = [], []
    for r in range(world):
        e, s_ = (em, src) if r == 0 else rank_inputs(r)
        e = e.copy()
        e["instance"] += r * V
        ems.append(e)
        srcs.append(s_)
    with orc.OracleMixer(**cfg_all) as o:
        setup_mixer(o, dict(w, voices=V * world), abi, [np.concatenate(ems)], listeners, areas)
        want, _ = o.mix_block(synth.make_voices(V * world), np.concatenate(srcs), F, want_peaks=False, threads=host_threads())
    ok, worst, nbad = S.sample_close(got, want)
    return {"ranks": world, "voices_total": V * world, "routing_exact": bool(np.array_equal(S.routing(got), S.routing(want))),
            "reduced_sum_within_1e-5_rel_or_-110dBFS_of_f32_oracle": ok, "samples_out": nbad, "worst_abs_err": worst,
            "peak_abs_bus_sample": float(np.abs(want).max())}


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 flop at the measured max SM clock


def flops_per_voice_frame(kind, C, n_send, stages=0, n_fx=0):
    """fp32 operations per voice-frame of the filtered paths, counted from the reference loops:
    ramp 4 + multiply 1 + biquad 9 + coefficient increments 5 per stream, plus 5 per (send, pair, side) for the bus ramp."""
    if kind == "B-filter":      # mix_channel: 2C streams of (ramp, mul, interpolated biquad), audio_spatializer_3d.cpp:589-597
        return 2 * C * (4 + 1 + 9 + 5) + 2 * C * n_send * 5
    if kind == "A-filter":      # process_frames: 2 interpolated biquads, then Q15 sends, :524-529
        return 2 * (9 + 5) + 2 * C * n_send * 5
    if kind == "effect":        # n_fx effects x stages biquads per side (no interpolation), audio_spatializer_effect.cpp:52-75
        return 2 * 9 * stages * n_fx + 2 * C * n_send * 5
    return 2 * C * n_send * 2   # unfiltered polynomial rows


EXTRA_CONFIGS = [
    dict(name="configs[1]: AudioSpatializer3D, 1024 voices, 5.1, inverse-square attenuation + attenuation filter, Mode B", voices=1024, frames=512,
         speaker_mode=2, num_buses=2, kind="B-filter", n_send=1.25,
         sc=dict(spat=dict(mix_channel_mode=1, attenuation_model=1), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25)),
    dict(name="configs[3]: AudioSpatializerEffect, 4096 voices, stereo, 1-stage high-shelf chain, Master/area bus + reverb bus", voices=4096,
         frames=512, speaker_mode=0, num_buses=3, kind="effect", stages=1, n_fx=1, n_send=1.5,
         sc=dict(effect_chain=[dict(mode=7, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=1)], effect_gain_binding=0,
                 area=dict(reverb_bus=2, amount=0.4, override_bus=True, bus=1), area_fraction=0.5)),
    dict(name="configs[3]: AudioSpatializerEffect, 4096 voices, stereo, 4-stage high-shelf chain, Master/area bus + reverb bus", voices=4096,
         frames=512, speaker_mode=0, num_buses=3, kind="effect", stages=4, n_fx=1, n_send=1.5,
         sc=dict(effect_chain=[dict(mode=7, cutoff_hz=4000.0, resonance=1.0, gain=0.3, stages=4)], effect_gain_binding=0,
                 area=dict(reverb_bus=2, amount=0.4, override_bus=True, bus=1), area_fraction=0.5)),
    dict(name="configs[2] with the attenuation filter ON, Mode A (process_frames: 2 biquads per voice-frame)", voices=16384, frames=512,
         speaker_mode=3, num_buses=2, kind="A-filter", n_send=1.25,
         sc=dict(spat=dict(mix_channel_mode=0), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25)),
    dict(name="configs[2] with the attenuation filter ON, Mode B (mix_channel: 8 biquads per voice-frame)", voices=16384, frames=512,
         speaker_mode=3, num_buses=2, kind="B-filter", n_send=1.25,
         sc=dict(spat=dict(mix_channel_mode=1), area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25), steps=48),
    dict(name="configs[4] corner: 256 voices x 128-frame blocks, stereo, filter off", voices=256, frames=128, speaker_mode=0, num_buses=2,
         kind="stream", n_send=1.25, sc=dict(spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0),
                                             area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, r_min=10.0)),
    dict(name="configs[4] corner: 65536 voices x 2048-frame blocks, 7.1, filter off", voices=65536, frames=2048, speaker_mode=3, num_buses=2,
         kind="stream", n_send=1.25, sc=dict(spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0),
                                             area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, r_min=10.0), sets=2, steps=12),
    dict(name="configs[4] corner: 4096 voices x 1024-frame blocks, 3.1, filter off", voices=4096, frames=1024, speaker_mode=1, num_buses=2,
         kind="stream", n_send=1.25, sc=dict(spat=dict(mix_channel_mode=1, unit_size=1.0, attenuation_filter_db=-80.0),
                                             area=dict(reverb_bus=1, amount=0.5), area_fraction=0.25, r_min=10.0)),
]


def measure_config(gas, torch, spec, device, peak_gbs, cpu_budget_s=2.5):
    """One secondary configuration on one GPU: us per block (CUDA-graph replay, CUDA events), fractions of the HBM and fp32
    rooflines, the CPU oracle beside it and a full-size parity check of two blocks."""
    import scenarios as S
    from oracle import orc
    abi, synth = gas.abi, gas.synth
    V, F, mode, B = spec["voices"], spec["frames"], spec["speaker_mode"], spec["num_buses"]
    C = mode + 1
    kw = dict(spec["sc"])
    r_min = kw.pop("r_min", 0.5)
    sc = S.default_scenario(voices=V, frames=F, speaker_mode=mode, num_buses=B, **kw)
    sets, steps = spec.get("sets", 4), spec.get("steps", 100)
    dev = torch.device("cuda", device)
    inst = np.arange(V, dtype=np.int32)
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(**sc["area"])], dtype=abi.area) if sc["area"] else None
    dt = F / sc["mix_rate"]
    cfg = dict(max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=B, speaker_mode=mode, mix_rate=sc["mix_rate"])
    ems = [synth.make_emitters(V, block=b, dt=dt, area_fraction=sc["area_fraction"], r_min=r_min) for b in range(sets)]
    voices = synth.make_voices(V)
    spat = S.make_spatializer(sc)

    def setup(mm):
        mm.spatializer_set(0, spat)
        mm.instance_init(inst, 0)
        mm.gain_compute(ems[0], listeners, areas, want_params=False)
        mm.instance_start(inst)
        mm.voice_init(inst)

    with gas.Mixer(device=device, **cfg) as m:
        setup(m)
        m.listeners_set(listeners)
        if areas is not None:
            m.areas_set(areas)
        d_em = [torch.from_numpy(e.view(np.uint8).copy()).to(dev) for e in ems]
        d_voices = torch.from_numpy(voices.view(np.uint8).copy()).to(dev)
        d_src = [(torch.rand((V, F, 2), device=dev) - 0.5) * 0.5 for _ in range(sets)]
        d_bus = [torch.zeros((B, C, F, 2), device=dev) for _ in range(NB)]
        n_graphs = sets * NB // math.gcd(sets, NB)

        def nxt(k):
            return dict(n_emitters=V, d_emitters=d_em[k % sets].data_ptr(), n_voices=V, d_voices=d_voices.data_ptr(), src_rows=V, frames=F,
                        d_bus_out=d_bus[k % NB].data_ptr())

        def capture():
            gs = []
            if CLASSIC:
                for s_ in range(n_graphs):
                    m.capture_begin()
                    m.mix_block_device(V, d_voices.data_ptr(), d_src[s_ % sets].data_ptr(), V, F, F, d_bus[s_ % NB].data_ptr())
                    m.gain_compute_device(V, d_em[(s_ + 1) % sets].data_ptr())
                    gs.append(m.capture_end())
                return gs
            m.step_device(d_src[0].data_ptr(), F, next=None)  # ends a run in progress (nothing happens when none is)
            m.step_device(next=nxt(0))
            for s_ in range(n_graphs):
                m.capture_begin()
                m.step_device(d_src[s_ % sets].data_ptr(), F, next=nxt(s_ + 1))
                gs.append(m.capture_end())
            return gs

        graphs = capture()
        stream = torch.cuda.ExternalStream(m.mix_stream, device=dev)
        for k in range(2 * n_graphs):
            m.graph_launch(graphs[k % n_graphs])
        m.step_join_device()
        m.sync()
        steps = (steps // n_graphs) * n_graphs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for k in range(steps):
            m.graph_launch(graphs[k % n_graphs])
        m.step_join_device()
        with torch.cuda.stream(stream):
            e1.record()
        m.sync()
        us = 1e3 * e0.elapsed_time(e1) / steps
        m.profile_enable(True)
        pg = capture()
        for k in range(4 * n_graphs):
            m.graph_launch(pg[k % n_graphs])
        prof = m.profile_read()
        m.profile_enable(False)
        if not CLASSIC:
            m.step_device(d_src[0].data_ptr(), F, next=None)
        m.sync()
        del d_src
    # parity: two state-carrying blocks, full size, against the oracle (bounded: the largest corner is checked on a slice of time)
    parity = None
    if V * F <= 16384 * 512:
        src_h = [synth.make_sources(V, F, block=b, mix_rate=sc["mix_rate"]) for b in range(2)]
        with gas.Mixer(device=device, **cfg) as m, orc.OracleMixer(**cfg) as o:
            ok_all, worst_all, routing = True, 0.0, True
            for mm in (m, o):
                setup(mm)
            for b in range(2):
                for mm in (m, o):
                    mm.gain_compute(ems[b % sets], listeners, areas, want_params=False)
                got, _ = m.mix_block(voices, src_h[b], F, want_peaks=False)
                want, _ = o.mix_block(voices, src_h[b], F, want_peaks=False, threads=host_threads())
                ok, worst, _ = S.sample_close(got, want)
                ok_all, worst_all = ok_all and ok, max(worst_all, worst)
                routing = routing and bool(np.array_equal(S.routing(got), S.routing(want)))
            parity = {"blocks": 2, "routing_exact": routing, "within_tolerance": ok_all, "worst_abs_err": worst_all}
    # CPU oracle beside it (bounded sample)
    cpu = None
    src0 = synth.make_sources(min(V, 4096), F, block=0, mix_rate=sc["mix_rate"])
    if V > 4096:
        src0 = np.tile(src0, (V // 4096, 1, 1))
    with orc.OracleMixer(**cfg) as o:
        setup(o)
        thr = host_threads()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < cpu_budget_s or n < 1:
            o.gain_compute(ems[n % sets], listeners, areas, want_params=False)
            o.mix_block(voices, src0, F, want_peaks=False, threads=thr)
            n += 1
        cs = time.perf_counter() - t0
        cpu = {"value": V * F * n / cs, "unit": UNIT, "cores": thr, "kind": "port", "sample": f"{n} blocks in {cs:.1f} s"}
    filt = spec["kind"] != "stream"
    bytes_blk = algorithmic_bytes(V, F, C, B, filter_on=(spec["kind"] == "B-filter"))
    fl = flops_per_voice_frame(spec["kind"], C, spec["n_send"], spec.get("stages", 0), spec.get("n_fx", 0)) * V * F
    return {"config": spec["name"], "voices": V, "frames": F, "channel_pairs": C, "buses": B, "us_per_block": us,
            "voice_frames_per_s": V * F / (us * 1e-6), "x_realtime": dt / (us * 1e-6),
            "roofline": {"hbm_frac": bytes_blk / (us * 1e-6) / 1e9 / peak_gbs, "algorithmic_bytes_per_block": bytes_blk,
                         "fp32_frac": (fl / (us * 1e-6) / 1e12 / FP32_PEAK_TFLOPS) if filt else None,
                         "fp32_flop_per_block": fl if filt else None, "fp32_peak_tflops": FP32_PEAK_TFLOPS,
                         "bound": "fp32 issue / recurrence latency" if filt else "hbm (launch latency for the small corner)"},
            "kernels_us": {k: 1e3 * v[0] / max(1, v[1]) for k, v in prof.items()}, "cpu_baseline": cpu, "parity": parity}


def k2_debug_dump(m, torch):
    """experiments (GAS_K2_DEBUG & 8): K2's in-kernel timeline (globaltimer stamps of every CTA) of the last replayed step, to stderr"""
    import ctypes
    torch.cuda.synchronize()
    keys = (ctypes.c_uint64 * 128)()
    counts = (ctypes.c_int32 * 256)()
    if m._lib.gas_debug_classes(m._ctx, keys, counts) == 0:
        for i in range(128):
            k = keys[i]
            if k and (counts[i] or counts[128 + i]):
                print(f"  class slot {i}: path {k & 3} mode {(k >> 2) & 3} flags {(k >> 4) & 0xf:#x} sends {(k >> 8) & 0xf} mask {(k >> 16) & 0xffff:#x} "
                      f"quad {(k >> 32) & 0xfff:#x} counts {counts[i]}/{counts[128 + i]}", file=sys.stderr)
    fn = m._lib.gas_debug_timeline
    fn.restype = ctypes.c_void_p
    fn.argtypes = [ctypes.c_void_p]
    ptr = fn(m._ctx)
    if ptr:
        tl = torch.empty(148 * 32, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(tl.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(148 * 32 * 8), 3)
        t = tl.cpu().numpy().reshape(148, 32).astype(np.float64)
        t0 = t[:, 0][t[:, 0] > 0].min()
        names = {0: "start", 11: "table in smem", 1: "partition", 12: "first indices", 13: "stage 0 issued", 2: "first data", 3: "last data",
                 9: "flush begins", 4: "flushed", 16: "control: entry", 23: "gain: emitter loaded", 24: "gain: tables loaded", 25: "gain: attenuation done", 26: "gain: pan done",
                 27: "gain: stores issued", 17: "control: gains done", 28: "plan: pass table reset", 29: "plan: voice computed",
                 30: "plan: registered in CTA table", 31: "plan: class ranges reserved", 18: "control: barrier passed",
                 19: "control: instances done", 20: "control: voices done", 21: "control: ticket", 22: "control: plan published"}
        for k, nm in names.items():
            v = (t[:, k][t[:, k] > 0] - t0) * 1e-3
            if v.size:
                print(f"  K2 timeline {nm:16s} min {v.min():7.2f} avg {v.mean():7.2f} max {v.max():7.2f} us", file=sys.stderr)
        units = t[:, 5]
        per = (t[:, 3] - t[:, 2]) * 1e-3 / np.maximum(units, 1)
        for u in sorted(set(units.astype(int))):
            sel = units == u
            print(f"  K2 units {u:2d}: n={int(sel.sum()):3d} per-unit {per[sel].mean():.3f} us  last data {((t[sel, 3] - t0) * 1e-3).mean():.2f}  "
                  f"flushed {((t[sel, 4] - t0) * 1e-3).mean():.2f}", file=sys.stderr)


def file_sha16(path):
    try:
        with open(path, "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except Exception:
        return None


def gpu_arm(args):
    import torch
    import gaspkg
    gas = gaspkg.load()
    abi, synth = gas.abi, gas.synth
    rank, local_rank, world = dist_env()
    heartbeat("imported")
    torch.cuda.set_device(local_rank)
    select_form(world)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = dict(WORKLOAD)
    if args.voices:
        w["voices"] = args.voices
    if args.frames:
        w["frames"] = args.frames
    if args.area_fraction is not None:
        w["area_fraction"] = args.area_fraction
    V, F, C, B = w["voices"], w["frames"], w["speaker_mode"] + 1, w["num_buses"]
    # every step graph is launched at least once before the timed region (first launches upload the graph); the warm-up is
    # rounded up to whole graphs
    K = args.steps
    chunk = steps_per_graph(K)
    W = max(3, args.warmup, N_SETS, graphs_per_cycle(chunk) * chunk)  # every graph is launched (uploaded) before the timed region
    W = ((W + chunk - 1) // chunk) * chunk

    parity_src = synth.make_sources(V, F, block=0, mix_rate=w["mix_rate"]) if (rank == 0 and not args.no_parity) else None
    heartbeat("process group")
    dw = DeviceWorkload(gas, torch, w, local_rank, rank, parity_src)
    m = dw.mixer
    heartbeat("workload resident")
    stream = torch.cuda.ExternalStream(m.mix_stream, device=torch.device("cuda", local_rank))

    parity = None
    host_inputs = None
    if rank == 0 and not args.no_parity:
        host_inputs = build_host_inputs(w, abi, synth)
    # (the parity gate itself runs after the timed passes: a second context that allocates and frees ~1.5 GB next to the
    # benchmarked one left the step 1.4 us slower for the rest of the process when it ran first)

    peer = dist is not None and args.reduce == "peer"
    if peer:
        dw.comm_setup(dist)
    if dist is not None and not peer:
        chunk = 1  # the NCCL variant reduces between graph launches
        W = max(3, args.warmup, N_SETS)
    heartbeat("exchange open")
    graphs = dw.capture_steps(chunk)
    heartbeat("captured")
    sampler = ClockSampler(local_rank)  # NVML initialised here, outside the timed region

    def run_steps(k0, n):
        # steps k0 .. k0 + n - 1 (k0 and n multiples of `chunk`): one graph launch per `chunk` steps
        for k in range(k0, k0 + n, chunk):
            m.graph_launch(graphs[(k // chunk) % len(graphs)])
            if dist is not None and not peer and not os.environ.get("GAS_BENCH_NOREDUCE"):
                with torch.cuda.stream(stream):
                    m.step_join_device()
                    dist.all_reduce(dw.d_bus[k % NB])  # sum of the per-GPU partial bus buffers (NCCL over NVLink)

    def barrier():
        # Drain the GPU first, THEN meet the other ranks: a collective's kernel enqueued while step kernels are in flight can take
        # an SM that a step launch needs for its 148th CTA (its control warps wait for every CTA of the launch), while the step
        # kernels keep the collective's other CTAs off the SMs: the 4- and 8-GPU runs of this round hung exactly there.
        m.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: `value` ----------------------------------------------------------------------------
    dw.reduce_prime()
    run_steps(0, W)
    dw.reduce_drain(W - 1)
    barrier()
    heartbeat("warm")
    sampler.start()
    if peer:
        # device-side start gate: an in-order reduce of a scratch buffer is one arrival round over peer memory, so every
        # rank's mix stream leaves it within an NVLink round trip of the others; host-side skew after the barrier above
        # stays outside the timed region
        m.reduce_bus_device(dw.d_gate.data_ptr(), F)
    launches0 = m.kernel_launches
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_wall0 = time.time()
    with torch.cuda.stream(stream):
        ev0.record()
    run_steps(W, K)
    with torch.cuda.stream(stream):
        ev1.record()
    gpu_launches = m.kernel_launches - launches0
    dw.reduce_drain(W + K - 1)  # the two blocks still in flight in the exchange pipeline: reported separately
    with torch.cuda.stream(stream):
        ev2.record()
    barrier()
    heartbeat("timed")
    t_wall1 = time.time()
    sampler.stop()
    sampler.join()
    ms = ev0.elapsed_time(ev1)
    drain_ms = ev1.elapsed_time(ev2)
    if dist is not None:
        t = torch.tensor([ms, drain_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, drain_ms = float(t[0].item()), float(t[1].item())
    value = world * V * F * K / (ms * 1e-3)
    clocks = sampler.summary(t_wall0, t_wall1)

    if os.environ.get("GAS_K2_DEBUG") and int(os.environ["GAS_K2_DEBUG"]) & 8:
        k2_debug_dump(m, torch)

    # ---- roofline: per-launch duration of the streaming mix kernel ----------------------------------------------
    # The same steps are captured once more with per-kernel timing on: the graphs then carry event-record nodes
    # around every kernel, so each duration is measured on the device, on the launching stream, as the kernel
    # runs inside the replayed step (timing every launch from the host would measure the host's launch rate).
    m.profile_enable(True)
    pgraphs = dw.capture_steps()
    kp = max(8, min(K, 256))
    dw.reduce_prime()
    for k in range(kp):
        m.graph_launch(pgraphs[k % N_SETS])
    dw.reduce_drain(kp - 1)
    prof = m.profile_read()
    # once more without K1 beside the mix: K2 with the GPU to itself (reported next to the in-step figure, not instead of it)
    dw.no_gain = True
    m.profile_enable(True)
    agraphs = dw.capture_steps()
    dw.reduce_prime()
    for k in range(kp):
        m.graph_launch(agraphs[k % N_SETS])
    dw.reduce_drain(kp - 1)
    prof_alone = m.profile_read()
    dw.no_gain = False
    m.profile_enable(False)
    barrier()
    heartbeat("profiled")
    parity_multi = None
    if dist is not None and peer and not args.no_parity:
        parity_multi = parity_gate_multi(gas, torch, dist, w, local_rank, rank, world, abi, synth)
    if rank == 0 and not args.no_parity:
        parity = parity_gate(gas, w, local_rank, abi, synth, parity_src, host_inputs)
        if parity_multi is not None:
            parity["multi_gpu_reduced_sum"] = parity_multi
    heartbeat("parity")
    k2_ms, k2_n = prof["mix_stream"]
    peak, peak_src = measured_hbm_peak()
    bytes_launch = algorithmic_bytes(V, F, C, B)
    k2_us = max(1e3 * k2_ms / max(1, k2_n), 1e-6)
    step_us = 1e3 * ms / K

    def us(kind):
        return 1e3 * prof[kind][0] / max(1, prof[kind][1])

    # The dominant kernel is the step kernel (one launch per step: streaming of block k + gains and plan of block k + 1 on its
    # control warps).  Consecutive launches overlap (programmatic dependent launch), so its average launch duration over the
    # timed region is the timed region divided by its launches; the isolated reading (event-record nodes around the kernel in
    # a second, profiled capture, which serialise the launches) is reported beside it.
    # (block-call form: the step is four kernels, so the streaming kernel's own duration is the isolated reading)
    launch_us = k2_us if CLASSIC else step_us
    achieved = bytes_launch / (launch_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "k_step (streaming of block k + gains and plan of block k+1)" if not CLASSIC else
                          "k_step without control work (block-call form: gains, plan and streaming are separate kernels)",
                "us_per_launch": launch_us, "algorithmic_bytes_per_launch": bytes_launch, "launches_in_timed_region": K,
                "peak_source": peak_src,
                "timing": ("CUDA events on the mix stream around the timed region / launches of the kernel (launches overlap by "
                           "programmatic dependent launch, so the per-launch average IS the step time)" if not CLASSIC else
                           "CUDA event-record nodes around the kernel inside a profiled capture of the same steps"),
                "step_frac_of_hbm_peak": bytes_launch / (step_us * 1e-6) / 1e9 / peak,
                "isolated": {"note": "same kernel between event-record nodes in a profiled capture of the same steps (the nodes serialise "
                                     "the launches and read ~2.7 us by themselves: event_pair_overhead_us)",
                             "us_per_launch": k2_us, "frac": bytes_launch / (k2_us * 1e-6) / 1e9 / peak,
                             "event_pair_overhead_us": 1e3 * prof["none"][0] / max(1, prof["none"][1]) if prof["none"][1] else None,
                             "without_gains_us_per_launch": 1e3 * prof_alone["mix_stream"][0] / max(1, prof_alone["mix_stream"][1])},
                "other_kernels_us": {"gain_K1": us("gain"), "plan": us("prologue"), "mix_voice_K3": us("mix_voice")}}
    # DRAM traffic of the kernel comes from one `ncu --set full` capture (profiles/): valid only for the kernel source it was
    # taken from, so the file carries the hash of gas_mix_stream.cu and a stale file reports null instead of an old number
    traffic_file = os.path.join(ROOT, "profiles", "k2_traffic_bytes.json")
    k2_src = os.path.join(ROOT, "godot-audio-spatializer_b200", "csrc", "gas_mix_stream.cu")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            if tj.get("k2_source_sha16") == file_sha16(k2_src):
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tj.get("source")
            else:
                roofline["traffic_note"] = "profiles/k2_traffic_bytes.json was captured from another version of gas_mix_stream.cu"
        except Exception:
            pass

    # ---- e2e: host buffers through the public API ---------------------------------------------------------------------
    if not CLASSIC:
        m.step_device(dw.d_src[0].data_ptr(), F, next=None)  # the pipelined run ends here: the block calls below plan for themselves
        m.sync()
    ke = max(4, min(K, args.e2e_steps))
    pin_src = [torch.empty((V, F, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    for t_ in pin_src:
        t_.copy_(dw.d_src[0].cpu() if parity_src is None else torch.from_numpy(parity_src))
    pin_voices = torch.from_numpy(dw.voices_host.view(np.uint8).copy()).pin_memory()
    pin_bus = torch.empty((B, C, F, 2), dtype=torch.float32).pin_memory()
    h2d = V * F * 8 + dw.voices_host.nbytes + dw.emitters_host[0].nbytes + dw.listeners.nbytes + dw.areas.nbytes
    d2h = B * C * F * 8
    d_e2e_src = torch.empty((V, F, 2), dtype=torch.float32, device=torch.device("cuda", local_rank)) if dist is not None else None

    def e2e_step(k):
        m.gain_compute(dw.emitters_host[k % N_SETS], dw.listeners, dw.areas, want_params=False)
        if dist is None:
            # gas_mix_block: sources and voices host -> device, the bus buffers device -> host, all inside the call
            m.mix_block_host_ptr(V, pin_voices.data_ptr(), pin_src[k % 2].data_ptr(), V, F, pin_bus.data_ptr())
        else:
            # N > 1: the rank's sources go up, its partial bus buffers are summed over peer memory (gas_reduce_bus_device)
            # and the complete sum comes back to the host
            with torch.cuda.stream(stream):
                d_e2e_src.copy_(pin_src[k % 2], non_blocking=True)
                dw.d_voices.copy_(pin_voices, non_blocking=True)
            m.mix_block_device(V, dw.d_voices.data_ptr(), d_e2e_src.data_ptr(), V, F, F, dw.d_bus[0].data_ptr())
            if peer:
                m.reduce_bus_device(dw.d_bus[0].data_ptr(), F)
            else:
                with torch.cuda.stream(stream):
                    dist.all_reduce(dw.d_bus[0])
            with torch.cuda.stream(stream):
                pin_bus.copy_(dw.d_bus[0], non_blocking=True)
            m.sync()

    for k in range(3):
        e2e_step(k)
    barrier()
    heartbeat("e2e warm")
    t0 = time.perf_counter()
    for k in range(ke):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * V * F * ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
           "ms_per_step": 1e3 * e2e_s / ke,
           "api": "gas_gain_compute + gas_mix_block (host pointers, pinned)" if dist is None else
                  "gas_gain_compute + pinned host->device copy + gas_mix_block_device + gas_reduce_bus_device + device->host copy of the sum"}

    # ---- e2e with device-resident sources (N = 1): PCM clips live in HBM, the resampler in front of the path produces the block's
    # source rows on the device (gas_mix_block_resident), so per step only the emitters and the voice list travel up and the bus
    # buffers come back.  The leg runs in a process of its own (resident_leg below) and is adopted only if its self-checks pass:
    # this code path was finished after the round's GPU budget was spent, so nothing that goes wrong in it may touch the line.
    if dist is None and not os.environ.get("GAS_BENCH_NO_RESIDENT"):
        import subprocess
        res = None
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--resident-leg", "--e2e-steps", str(max(8, min(K, args.e2e_steps))),
                                 "--voices", str(args.voices), "--frames", str(args.frames)], capture_output=True, text=True, timeout=300)
            for ln in reversed(cp.stdout.strip().splitlines()):
                if ln.startswith("{"):
                    res = json.loads(ln)
                    break
            if res is None:
                res = {"error": f"exit {cp.returncode}: {cp.stderr.strip()[-300:]}"}
        except Exception as ex:
            res = {"error": repr(ex)[:300]}
        if res.get("value") and res.get("parity_ok") and res.get("all_voices_still_active") and res.get("bus_finite_and_nonzero"):
            e2e = dict(res, host_frames=e2e)
        else:
            e2e["resident_sources"] = res

    # ---- cpu baseline (rank 0, N = 1) --------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        inputs = host_inputs or build_host_inputs(w, abi, synth)
        v_all, n_all, s_all = run_cpu(w, 400, 1, threads, abi, synth, budget_s=12.0, inputs=inputs)
        v_one, n_one, s_one = run_cpu(w, 400, 1, 1, abi, synth, budget_s=8.0, inputs=inputs)
        cpu = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_all} full blocks ({V} voices x {F} frames, gain + mix) in {s_all:.1f} s, OpenMP over instances",
               "single_thread": {"value": v_one, "blocks": n_one, "seconds": s_one,
                                 "note": "faithful: Godot mixes on one audio thread"},
               "note": "oracle port of the reference loop (gcc -O2), pinned bit for bit against the reference module's own code "
                       "(oracle/_ref); headless Godot cannot be built here"}

    # ---- the other BASELINE.json configurations (rank 0, N = 1, time-boxed) ---------------------------------------------------
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        dw.mixer.close()
        del dw
        torch.cuda.empty_cache()
        configs = []
        t_cfg0 = time.perf_counter()
        for spec in EXTRA_CONFIGS:
            if time.perf_counter() - t_cfg0 > args.configs_budget:
                configs.append({"config": spec["name"], "skipped": "time box"})
                continue
            try:
                configs.append(measure_config(gas, torch, spec, local_rank, peak))
            except Exception as ex:  # a secondary configuration must not take the headline line down with it
                configs.append({"config": spec["name"], "error": repr(ex)[:300]})

    heartbeat("e2e")
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(w, world, peer),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clocks,
            "parity": parity, "pdl": os.environ.get("GAS_PDL", "default"), "steps_per_graph_launch": chunk,
            "exchange_drain_ms": drain_ms if world > 1 else None, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        # orderly exit at N > 1: every rank leaves together, and interpreter teardown (tensors freed after the mixer's streams
        # and the process group are gone) is skipped
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()
        os._exit(0)


def resident_leg(args):
    """Child process of the N = 1 bench: e2e with device-resident sources.  Prints one JSON object."""
    import torch
    import gaspkg
    import scenarios as S
    from oracle import orc
    gas = gaspkg.load()
    abi, synth = gas.abi, gas.synth
    w = dict(WORKLOAD)
    if args.voices:
        w["voices"] = args.voices
    if args.frames:
        w["frames"] = args.frames
    V, F, C, B = w["voices"], w["frames"], w["speaker_mode"] + 1, w["num_buses"]
    out = {"unit": UNIT}

    def clip(n, seed, rate=44100.0):
        rs = np.random.RandomState(seed)
        tt = np.arange(n, dtype=np.float64)
        return (0.25 * np.sin(2 * np.pi * (80.0 + 23.0 * seed) * tt / rate)[:, None] + 0.02 * rs.randn(n, 2)).astype(np.float32)

    # ---- self-check: 96 voices, 6 blocks, clips that end inside the run: the same calls on the CUDA mixer and on the oracle twin
    # (upstream's resampler per voice + the stream form of the mix), compared block by block
    try:
        Vc, nblk = 96, 6
        cfg = dict(max_instances=Vc, max_voices=Vc, max_frames=F, num_buses=B, speaker_mode=w["speaker_mode"], mix_rate=w["mix_rate"])
        listeners = np.array([abi.identity_listener()], dtype=abi.listener)
        areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
        clips = [clip(1400 + 173 * k_, 10 + k_) for k_ in range(6)]
        inst = np.arange(Vc, dtype=np.int32)
        voices = synth.make_voices(Vc)
        ems = [synth.make_emitters(Vc, block=b_, dt=F / w["mix_rate"], area_fraction=0.5) for b_ in range(nblk)]
        for e_ in ems:
            e_["pitch_scale"] = np.linspace(0.5, 2.0, Vc).astype(np.float32)

        def play(mm):
            mm.spatializer_set(0, abi.spatializer_defaults(**w["spat"]))
            mm.instance_init(inst, 0)
            mm.gain_compute(ems[0], listeners, areas, want_params=False)
            mm.instance_start(inst)
            mm.voice_init(inst)
            for k_, c_ in enumerate(clips):
                mm.source_set(k_, c_, 44100.0)
            mm.voice_play(inst, inst % len(clips))
            active, res = np.ones(Vc, dtype=bool), []
            for b_ in range(nblk):
                mm.gain_compute(ems[b_], listeners, areas, want_params=False)
                live = voices[active].copy()
                live["src_row"] = np.arange(live.size)
                bus_, status_ = mm.mix_block_resident(live, F)
                res.append((bus_, status_.copy()))
                alive = (status_ & abi.VOICE_ACTIVE) != 0
                idx = np.nonzero(active)[0]
                active[idx[~alive]] = False
            return res

        with gas.Mixer(**cfg) as m, orc.OracleMixer(**cfg) as o:
            got, want = play(m), play(o)
        ok_all, worst_all = True, 0.0
        for (gb, gs), (wb, ws) in zip(got, want):
            same = gs.shape == ws.shape and bool(np.array_equal(gs, ws)) and gb.shape == wb.shape
            ok, worst, _ = S.sample_close(gb, wb) if same else (False, float("inf"), 0)
            ok_all = ok_all and same and ok and bool(np.array_equal(S.routing(gb), S.routing(wb)))
            worst_all = max(worst_all, worst)
        out["parity_ok"] = bool(ok_all)
        out["parity_worst_abs_err"] = worst_all
        out["parity_check"] = f"{Vc} voices x {nblk} blocks against the oracle twin (upstream resampler + stream form), clips ending inside the run"
    except Exception as ex:
        out["parity_ok"] = False
        out["parity_error"] = repr(ex)[:300]
        print(json.dumps(out), flush=True)
        return

    # ---- timed leg: the bench workload, 16 looping clips at 44.1 kHz ----------------------------------------------------
    listeners = np.array([abi.identity_listener()], dtype=abi.listener)
    areas = np.array([synth.reverb_area(reverb_bus=1, amount=0.5, uniformity=0.0)], dtype=abi.area)
    dt = F / w["mix_rate"]
    emitters = [synth.make_emitters(V, block=b_, dt=dt, area_fraction=w["area_fraction"], r_min=w["r_min"], r_max=w["r_max"]) for b_ in range(N_SETS)]
    voices = synth.make_voices(V)
    with gas.Mixer(device=0, max_instances=V, max_voices=V, max_frames=F, max_spatializers=2, num_buses=B, speaker_mode=w["speaker_mode"],
                   mix_rate=w["mix_rate"]) as m:
        setup_mixer(m, w, abi, emitters, listeners, areas)
        n_clips, clip_len = 16, 48000
        for c_ in range(n_clips):
            m.source_set(c_, clip(clip_len, c_), 44100.0, loop=True)
        vid = np.arange(V, dtype=np.int32)
        m.voice_play(vid, vid % n_clips, (vid * 977) % (clip_len - 256))
        pin_voices = torch.from_numpy(voices.view(np.uint8).copy()).pin_memory()
        pin_bus = torch.empty((B, C, F, 2), dtype=torch.float32).pin_memory()
        pin_status = torch.empty((V,), dtype=torch.int32).pin_memory()

        def step(k_):
            m.gain_compute(emitters[k_ % N_SETS], listeners, areas, want_params=False)
            m.mix_block_resident_host_ptr(V, pin_voices.data_ptr(), F, pin_bus.data_ptr(), pin_status.data_ptr())

        for k_ in range(4):
            step(k_)
        torch.cuda.synchronize()
        kr = max(8, args.e2e_steps)
        t0 = time.perf_counter()
        for k_ in range(kr):
            step(k_)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        bus = pin_bus.numpy()
        out.update({"value": V * F * kr / sec, "ms_per_step": 1e3 * sec / kr, "steps": kr,
                    "h2d_bytes_per_step": int(voices.nbytes + emitters[0].nbytes + listeners.nbytes + areas.nbytes),
                    "d2h_bytes_per_step": int(B * C * F * 8 + V * 4),
                    "api": "gas_gain_compute (host emitters) + gas_mix_block_resident (host voice list; resampler + voice lifecycle + mix on the "
                           "device; bus buffers and voice status back to the host)",
                    "sources": f"{n_clips} looping PCM clips of {clip_len} frames at 44.1 kHz resident in HBM (uploaded once with gas_source_set), "
                               "resampled to 48 kHz x pitch_scale per block",
                    "all_voices_still_active": bool((pin_status.numpy() & 1).all()),
                    "bus_finite_and_nonzero": bool(np.isfinite(bus).all() and np.abs(bus).max() > 0)})
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="N > 1: how the partial bus buffers are summed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE.json configurations")
    ap.add_argument("--configs-budget", type=float, default=130.0, help="seconds after which remaining secondary configurations are skipped")
    ap.add_argument("--area-fraction", type=float, default=None, help="fraction of voices inside the reverb area (experiments)")
    ap.add_argument("--resident-leg", action="store_true", help="internal: the e2e leg with device-resident sources, in a process of its own")
    args = ap.parse_args()
    if args.resident_leg:
        resident_leg(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        _, _, world = dist_env()
        if (world >= 4 and not os.environ.get("GAS_BENCH_CHILD") and not os.environ.get("GAS_BENCH_PIPELINED")
                and not os.environ.get("GAS_BENCH_CLASSIC") and not os.environ.get("GAS_BENCH_NO_ATTEMPT")):
            if guarded_pipelined_attempt(args):
                return
        gpu_arm(args)


if __name__ == "__main__":
    main()
