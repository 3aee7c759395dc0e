#!/bin/bash
# A/B of schedule knobs inside the bench step, on the GPU box:  bash tools/run_exp.sh  (edit the list below)
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for cfg in "X=0" "GAS_PDL=2" "GAS_K2_FIXED_COST=24" "GAS_K3_PARALLEL=1" "GAS_K1_SHAPE=1" "GAS_BENCH_NOGAIN=1"; do
  tag=$(echo $cfg | tr -cd 'A-Za-z0-9')
  timeout 120 bash tools/exp_bench.sh $tag "$cfg" > /dev/null
done
cat gpurun_out/exp_summary.txt
