#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
timeout 120 bash tools/exp_bench.sh fold "" > /dev/null
timeout 120 bash tools/exp_bench.sh fold2 "" > /dev/null
timeout 120 bash tools/exp_bench.sh rep1 "GAS_K2_REPLICAS=1" > /dev/null
echo "pytest: $(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)" >> gpurun_out/exp_summary.txt
