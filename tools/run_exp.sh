#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for f in 40 24 12 6 40; do
  timeout 120 bash tools/exp_bench.sh f$f "GAS_K2_FIXED_COST=$f" > /dev/null
done
timeout 120 bash tools/exp_bench.sh st4 "GAS_K2_STAGES=4" > /dev/null
timeout 120 bash tools/exp_bench.sh rep1 "GAS_K2_REPLICAS=1" > /dev/null
