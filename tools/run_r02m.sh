#!/bin/bash
# r02m: timeline with the evict-first hint, ring depth sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02m; mkdir -p $O
B="python bench.py --no-cpu --no-configs --no-parity --steps 64 --warmup 8 --e2e-steps 4"
for st in 6 4 3 2; do
GAS_K2_STAGES=$st GAS_K2_DEBUG=8 timeout 300 $B > $O/bench_tl_s$st.json 2> $O/bench_tl_s$st.err; echo "tl $st exit $?" >> $O/runs.log
GAS_K2_STAGES=$st timeout 300 $B > $O/bench_s$st.json 2> $O/bench_s$st.err; echo "s$st exit $?" >> $O/runs.log
GAS_K2_STAGES=$st GAS_PDL=12 timeout 300 $B > $O/bench_pdl_s$st.json 2> $O/bench_pdl_s$st.err; echo "pdl s$st exit $?" >> $O/runs.log
done
