#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
timeout 120 bash tools/exp_bench.sh noinl "" > /dev/null
timeout 120 bash tools/exp_bench.sh noinl_alone "GAS_SKIP=7" > /dev/null
timeout 120 bash tools/exp_bench.sh noinl2 "" > /dev/null
