#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02e; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
B="python bench.py --no-cpu --no-parity --no-configs --steps 320 --warmup 24 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run default GAS_DUMMY=1
run nogain GAS_BENCH_NOGAIN=1
run nogate GAS_K1_GATE=0
run pdl7 GAS_PDL=7
run pdl6 GAS_PDL=6
run timeline GAS_K2_DEBUG=8
env timeout 300 $B --area-fraction 0 > $O/bench_area0.json 2> $O/bench_area0.err
