#!/bin/bash
# r02l: in-kernel timeline of the step kernel (streaming + control warps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02l; mkdir -p $O
B="python bench.py --no-cpu --no-configs --no-parity --steps 64 --warmup 8 --e2e-steps 4"
GAS_K2_DEBUG=8 timeout 300 $B > $O/bench_tl.json 2> $O/bench_tl.err; echo "tl exit $?" >> $O/runs.log
GAS_K2_DEBUG=8 GAS_BENCH_NOGAIN=1 timeout 300 $B > $O/bench_tl_nogain.json 2> $O/bench_tl_nogain.err; echo "tl nogain exit $?" >> $O/runs.log
