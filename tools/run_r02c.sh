#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
B="python bench.py --no-cpu --no-parity --no-configs --steps 320 --warmup 24 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run default GAS_DUMMY=1
run chunk1 GAS_BENCH_CHUNK=1
run minb6 GAS_PROLOGUE_MINB=6
run minb7 GAS_PROLOGUE_MINB=7
run minb8 GAS_PROLOGUE_MINB=8
run minb7pdl7 GAS_PROLOGUE_MINB=7 GAS_PDL=7
run minb7pdl6 GAS_PROLOGUE_MINB=7 GAS_PDL=6
run minb7pdl5 GAS_PROLOGUE_MINB=7 GAS_PDL=5
run pdl0 GAS_PDL=0
run timeline GAS_K2_DEBUG=8 GAS_PROLOGUE_MINB=7
