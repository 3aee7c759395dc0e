#!/bin/bash
# r02j: bench with the pipelined step kernel; classic form beside it; K2 timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02j; mkdir -p $O
B="python bench.py --no-cpu --no-configs --steps 320 --warmup 24 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run pipelined GAS_DUMMY=1
run pdl8 GAS_PDL=12
run classic GAS_BENCH_CLASSIC=1
run nogain GAS_BENCH_NOGAIN=1
run nogain_pdl GAS_BENCH_NOGAIN=1 GAS_PDL=12
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs > $O/bench_driver.json 2> $O/bench_driver.err; echo "driver exit $?" >> $O/runs.log
