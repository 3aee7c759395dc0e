#!/bin/bash
# schedule-knob sweep of the step (bench.py) and of K2's in-kernel timeline (k2bench)
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for cfg in "GAS_K2_COST_FLOOR=0 GAS_K2_FIXED_COST=80" "GAS_K2_COST_FLOOR=26 GAS_K2_FIXED_COST=9" "GAS_K2_COST_FLOOR=26 GAS_K2_FIXED_COST=0" "GAS_K2_COST_FLOOR=26 GAS_K2_FIXED_COST=18" "GAS_K2_COST_FLOOR=20 GAS_K2_FIXED_COST=9" "GAS_K2_COST_FLOOR=32 GAS_K2_FIXED_COST=9"; do
  tag=$(echo $cfg | tr -d 'A-Z_ =' )
  timeout 120 bash tools/exp_bench.sh c$tag "$cfg"
  echo "## $cfg" >> gpurun_out/exp_summary.txt
  timeout 30 env $cfg GAS_K2_DEBUG=8 tools/k2bench 16384 512 0.25 16 | grep -E "last data|flushed" >> gpurun_out/exp_summary.txt
done
