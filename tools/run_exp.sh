#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
timeout 120 bash tools/exp_bench.sh base "" > /dev/null
timeout 120 bash tools/exp_bench.sh nogain "GAS_BENCH_NOGAIN=1" > /dev/null
timeout 120 bash tools/exp_bench.sh skipk3 "GAS_SKIP=4 GAS_K2_REPLICAS=1" > /dev/null
timeout 120 bash tools/exp_bench.sh skipk3ng "GAS_SKIP=4 GAS_K2_REPLICAS=1 GAS_BENCH_NOGAIN=1" > /dev/null
timeout 120 bash tools/exp_bench.sh onlypro "GAS_SKIP=6 GAS_BENCH_NOGAIN=1" > /dev/null
timeout 120 bash tools/exp_bench.sh onlyk2 "GAS_SKIP=5 GAS_BENCH_NOGAIN=1" > /dev/null
