#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
B="python bench.py --no-cpu --no-parity --no-configs --steps 300 --warmup 20 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run default GAS_DUMMY=1
run nogate GAS_K1_GATE=0
run k1s3 GAS_K1_SHAPE=3
run k1s4 GAS_K1_SHAPE=4
run k1s5 GAS_K1_SHAPE=5
run k1s6 GAS_K1_SHAPE=6
run k1s3pdl7 GAS_K1_SHAPE=3 GAS_PDL=7
run timeline GAS_K2_DEBUG=8
run timeline3 GAS_K2_DEBUG=8 GAS_K1_SHAPE=3
