#!/bin/bash
# r02q: fine-grained gain->plan dependency, early publish, raw-table start-up: tests, timeline, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "all exit $?" >> $O/runs.log
B="python bench.py --no-cpu --no-configs --no-parity --steps 64 --warmup 8 --e2e-steps 4"
GAS_K2_DEBUG=8 timeout 300 $B > $O/bench_tl.json 2> $O/bench_tl.err; echo "tl exit $?" >> $O/runs.log
timeout 300 $B > $O/bench.json 2> $O/bench.err; echo "b exit $?" >> $O/runs.log
GAS_PDL=4 timeout 300 $B > $O/bench_nopdl.json 2> $O/bench_nopdl.err; echo "nopdl exit $?" >> $O/runs.log
GAS_BENCH_CLASSIC=1 timeout 300 $B > $O/bench_classic.json 2> $O/bench_classic.err; echo "classic exit $?" >> $O/runs.log
