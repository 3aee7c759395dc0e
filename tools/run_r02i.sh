#!/bin/bash
# r02i: first GPU run of the fused step kernel: the new tests first (under a short timeout), then the whole gpu suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02i; mkdir -p $O
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > $O/pytest_parity.log 2>&1; echo "parity exit $?" >> $O/runs.log
timeout 300 python -m pytest tests/test_step_gpu.py -m gpu -q -x > $O/pytest_step.log 2>&1; echo "step exit $?" >> $O/runs.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "all exit $?" >> $O/runs.log
