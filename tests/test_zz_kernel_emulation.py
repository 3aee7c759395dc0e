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
