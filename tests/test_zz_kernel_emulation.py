"""The `gpu` test-suite executed WITHOUT a GPU: the product's .cu files (kernels and the C ABI's host code alike) compiled by g++
against tests/emu/ (a stand-in CUDA runtime: one fiber per CUDA thread, one OS thread per CTA, real atomics between CTAs, streams /
events / graph capture) and driven by the very same tests that run on the B200, each comparing with the oracle.

What this does and does not show: it executes every kernel's indexing, control flow, barriers, warp collectives, cross-CTA
protocols and the launch / capture plumbing around them; it says nothing about timing, PTX-level memory ordering or performance.
It caught real things when it was written (tests/emu/README.md).  The run happens in a child process (GAS_EMU=1 switches
tests/conftest.py to the emulation library) so that this process keeps the real binding.
"""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# every file of the gpu suite except the ones that need real peers / the compiled host mirror against a device
FILES = ["test_api_gpu.py", "test_golden.py", "test_lifecycle.py", "test_parity_gpu.py", "test_properties_gpu.py", "test_step_gpu.py",
         "test_zz_busgraph.py", "test_zz_graph_guard.py", "test_zz_resample.py"]


@pytest.mark.timeout(1500)
def test_gpu_suite_passes_on_the_emulated_device():
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    env = dict(os.environ, GAS_EMU="1", GAS_EMU_DEADLOCK_S="60")
    env.pop("PYTEST_CURRENT_TEST", None)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:faulthandler", "-p", "no:cacheprovider"] + [os.path.join(HERE, f) for f in FILES]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1400)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, "emulated gpu suite failed:\n" + tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail


def test_smoke_runs_on_the_emulated_device():
    """__graft_entry__.smoke() — what the driver runs first on the GPU box — against the emulation library."""
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    env = dict(os.environ, GAS_EMU="1", GAS_EMU_DEADLOCK_S="60")
    code = ("import sys, os; sys.path.insert(0, os.path.join(%r, 'tests'));"
            "import conftest; conftest._install_emulation();"
            "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SMOKE_OK" in r.stdout, (r.stdout + r.stderr)[-4000:]


def _emulated(cmd, timeout, ranks=1, port=29650, extra_env=None):
    env = dict(os.environ, GAS_EMU_DEADLOCK_S="120", GAS_EMU_SMS="4", **(extra_env or {}))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "GAS_EMU"):
        env.pop(k, None)
    launcher = os.path.join(HERE, "emu", "run_emulated.py")
    if ranks > 1:
        full = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
                "--master-port", str(port), launcher] + cmd
    else:
        full = [sys.executable, launcher] + cmd
    return subprocess.run(full, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def _bench_line(stdout):
    import json
    for ln in reversed(stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


BENCH_SMALL = ["--voices", "512", "--frames", "128", "--steps", "20", "--warmup", "5", "--e2e-steps", "4", "--no-configs"]


def test_bench_single_gpu_flow_on_the_emulated_device():
    """bench.py end to end at N = 1 (a small workload): graphs of the pipelined step kernel, profiled captures, the e2e legs (the
    resident-sources leg in its child process included), the CPU baseline, the full parity gate — every key of the JSON line."""
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    r = _emulated([os.path.join(ROOT, "bench.py")] + BENCH_SMALL, 900)
    line = _bench_line(r.stdout)
    assert r.returncode == 0 and line, (r.stdout + r.stderr)[-3000:]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
                "parity", "config"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 20 and line["gpu_launches"] == 40  # step kernel + voice-parallel kernel per step
    for blk in ("block0", "block1"):
        p = line["parity"][blk]
        assert p["routing_exact"] and p["within_1e-5_rel_or_-110dBFS_of_f32_oracle"]
        assert p["peak_abs_bus_sample"] > 0 and p["max_abs_err_gpu_vs_f64_shadow"] < 1e-5 * max(1.0, p["peak_abs_bus_sample"])
    # the resident-sources leg passed its own oracle check and was adopted
    assert line["e2e"].get("parity_ok") is True and "host_frames" in line["e2e"], line["e2e"]


@pytest.mark.parametrize("ranks,hang,form", [(2, False, "pipelined"), (4, False, "pipelined"), (4, True, "block-call")])
def test_bench_multi_gpu_flow_on_emulated_ranks(ranks, hang, form):
    """bench.py under torchrun with emulated ranks (one process each, exchange buffers mapped between the processes): the step
    graphs with the peer-memory exchange inside, the start gate, the drain, the N > 1 e2e leg and the reduced-sum parity gate.
    At 4 ranks the pipelined form runs in watchdog-guarded child processes (bench.guarded_pipelined_attempt); with `hang` the
    child of rank 1 stops after the warm-up, the watchdogs kill all children and the ranks measure with the block-call form."""
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    extra = dict(GAS_BENCH_TEST_HANG="warm", GAS_BENCH_HB_TIMEOUT="15") if hang else {}
    r = _emulated([os.path.join(ROOT, "bench.py"), "--gpus", str(ranks)] + BENCH_SMALL, 900, ranks=ranks, port=29650 + 3 * ranks + int(hang), extra_env=extra)
    line = _bench_line(r.stdout)
    assert r.returncode == 0 and line, (r.stdout + r.stderr)[-3000:]
    assert line["n_gpus"] == ranks
    pm = line["parity"]["multi_gpu_reduced_sum"]
    assert pm["ranks"] == ranks and pm["routing_exact"] and pm["reduced_sum_within_1e-5_rel_or_-110dBFS_of_f32_oracle"], pm
    assert ("gas_step_device" in line["config"]["launch"]) == (form == "pipelined")
    if ranks >= 4:
        assert ("form" in line) == (form == "pipelined")
        assert ("abandoned" in r.stderr) == hang


def test_peer_memory_reduce_on_emulated_ranks():
    """tools/check_multi_gpu.py (what tests/test_comm_gpu.py runs on real GPUs) with 4 emulated ranks."""
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    r = _emulated([os.path.join(ROOT, "tools", "check_multi_gpu.py")], 900, ranks=4, port=29661)
    assert r.returncode == 0 and "multi-GPU check ok" in r.stdout, (r.stdout + r.stderr)[-3000:]


@pytest.mark.parametrize("form", ["block-call", "pipelined"])
def test_randomised_parity_campaign_on_the_emulated_device(form):
    """tools/fuzz_parity.py: seeded random scenarios over every knob of the path (speaker modes, Mode A / B, effect chains, all
    attenuation models, areas, overriding buses, two listeners, Doppler, polyphony, late starts, silent rows, peaks, odd block sizes),
    CUDA sources on the emulated device against the oracle."""
    if not shutil.which(os.environ.get("CXX", "g++")):
        pytest.skip("needs g++")
    cmd = [os.path.join(ROOT, "tools", "fuzz_parity.py"), "--cases", "80", "--seed", "11"] + (["--pipelined"] if form == "pipelined" else [])
    r = _emulated(cmd, 900)
    assert r.returncode == 0 and "80 / 80 cases match the oracle" in r.stdout, (r.stdout + r.stderr)[-3000:]
