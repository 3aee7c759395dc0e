#!/bin/bash
# Round-2 record: the driver's bench line (N = 1), the default bench line, the reference arm, the ncu launch list of the same
# command (plain run first), and one `ncu --set full` capture of the step kernel (DRAM traffic, stalls).   bash tools/run_record2.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
tag=${1:-r02}
O=gpurun_out/$tag; mkdir -p $O
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; echo "driver bench rc=$?" >> $O/runs.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/runs.log
timeout 900 python bench.py --no-configs --no-cpu > $O/bench_default.json 2> $O/bench_default.err; echo "default bench rc=$?" >> $O/runs.log
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --no-configs --e2e-steps 4"
timeout 300 $CMD > $O/plain.log 2>&1; echo "plain rc=$?" >> $O/runs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/runs.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step' -s 20 -c 2 -f -o $O/k_step_full $CMD > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/runs.log
