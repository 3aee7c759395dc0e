#!/bin/bash
# round 2, first GPU pass: parity tests, then the step under a few scheduling knobs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
B="python bench.py --no-cpu --no-parity --no-configs --steps 300 --warmup 20 --e2e-steps 4"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name exit $?" >> $O/runs.log; }
run default GAS_DUMMY=1
run noscaled GAS_K2_SCALED=0
run pdl0 GAS_PDL=0
run pdl7 GAS_PDL=7
run pdl6 GAS_PDL=6
run r1like GAS_PDL=0 GAS_K2_REPLICAS=8 GAS_K2_SCALED=0
run timeline GAS_K2_DEBUG=8
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full exit $?" >> $O/runs.log
