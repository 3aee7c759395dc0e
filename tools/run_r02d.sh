#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02d; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
B="python bench.py --no-cpu --no-parity --no-configs --steps 320 --warmup 24 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run default GAS_PROLOGUE_MINB=7
run nogain GAS_PROLOGUE_MINB=7 GAS_BENCH_NOGAIN=1
run skip1 GAS_PROLOGUE_MINB=7 GAS_SKIP=1
run skip2 GAS_PROLOGUE_MINB=7 GAS_SKIP=2
run skip4 GAS_PROLOGUE_MINB=7 GAS_SKIP=4
run skip1nogain GAS_PROLOGUE_MINB=7 GAS_SKIP=1 GAS_BENCH_NOGAIN=1
run skip3nogain GAS_PROLOGUE_MINB=7 GAS_SKIP=3 GAS_BENCH_NOGAIN=1
run skip6nogain GAS_PROLOGUE_MINB=7 GAS_SKIP=6 GAS_BENCH_NOGAIN=1
run skip5nogain GAS_PROLOGUE_MINB=7 GAS_SKIP=5 GAS_BENCH_NOGAIN=1
run skip7 GAS_PROLOGUE_MINB=7 GAS_SKIP=7
