#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for v in 0 1 2 3 4 5 6 7 8 9; do
  timeout 120 bash tools/exp_bench.sh v$v "GAS_K1_VARIANT=$v" > /dev/null
done
timeout 120 bash tools/exp_bench.sh alone1 "GAS_K1_VARIANT=1 GAS_SKIP=7" > /dev/null
timeout 120 bash tools/exp_bench.sh alone2 "GAS_K1_VARIANT=2 GAS_SKIP=7" > /dev/null
