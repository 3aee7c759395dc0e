#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
timeout 120 bash tools/exp_bench.sh par1 "GAS_K3_PARALLEL=1" > /dev/null
timeout 120 bash tools/exp_bench.sh par0 "GAS_K3_PARALLEL=0" > /dev/null
timeout 120 bash tools/exp_bench.sh par1b "GAS_K3_PARALLEL=1" > /dev/null
timeout 120 bash tools/exp_bench.sh par0r1 "GAS_K3_PARALLEL=0 GAS_K2_REPLICAS=1" > /dev/null
