#!/bin/bash
# Round record, part 1: full bench line, reference arm, ncu launch list (same command, plain run first).
mkdir -p gpurun_out
tag=${1:-r01}
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --e2e-steps 4"
timeout 300 $CMD > gpurun_out/plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
echo "ncu launches rc=$?"
cat gpurun_out/bench_$tag.json | cut -c1-2400
