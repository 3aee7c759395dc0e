#!/bin/bash
# in-kernel timeline of the step kernel only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/tl; mkdir -p $O
B="python bench.py --no-cpu --no-configs --no-parity --steps 64 --warmup 8 --e2e-steps 4"
GAS_K2_DEBUG=8 timeout 300 $B > $O/bench_tl.json 2> $O/bench_tl.err; echo "tl exit $?" >> $O/runs.log
