"""Peer-memory bus reduce (csrc/gas_comm.cu) across GPUs: every rank mixes its shard, gas_reduce_bus_device (in order) and
gas_reduce_bus_exchange_device (one block in flight) sum the partial bus buffers over NVLink, every rank compares with the
oracle's unsharded mix (tools/check_multi_gpu.py under torchrun).  Skipped below two visible GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_reduced_bus_buffers_match_unsharded_oracle(ranks):
    if _gpus() < ranks:
        pytest.skip(f"needs {ranks} visible GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + ranks), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU check ok" in out.stdout, out.stdout[-2000:]
