#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/k2bench.txt; : > $out
run() { echo "## $*" >> $out; timeout 30 env "$@" >> $out 2>&1 || echo "   (exit $?)" >> $out; }
B="stdbuf -o0 tools/k2bench"
for f in 8 16 24 40; do run GAS_K2_DEBUG=8 GAS_K2_FIXED_COST=$f $B 16384 512 0.25 16; done
