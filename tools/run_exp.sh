#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for p in 0 2 0 2 6 0 2; do
timeout 120 bash tools/exp_bench.sh pdl$p "GAS_PDL=$p" > /dev/null
done
