#!/bin/bash
# r02g: baseline of this session: gpu tests, the driver's bench line, reference arm, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02g; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 exit $?" >> $O/runs.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?" >> $O/runs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-parity --no-configs --steps 4 --warmup 3 --e2e-steps 1 > $O/ncu.log 2>&1; echo "ncu exit $?" >> $O/runs.log
