#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/k2bench.txt; : > $out
run() { echo "## $*" >> $out; timeout 30 env "$@" >> $out 2>&1 || echo "   (exit $?)" >> $out; }
B=tools/k2bench
run X=1 $B 16384 512 0.25 16
run GAS_K2_DEBUG=8 $B 16384 512 0.25 16
run GAS_K2_DEBUG=11 $B 16384 512 0.25 16
run GAS_K2_DEBUG=8 GAS_K2_FIXED_COST=200 $B 16384 512 0.25 16
run GAS_K2_DEBUG=8 $B 16384 512 0 16
run GAS_K2_DEBUG=8 $B 2048 512 0.25 16
