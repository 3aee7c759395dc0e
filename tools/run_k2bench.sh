#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/k2bench.txt; : > $out
run() { echo "## $*" >> $out; timeout 30 env "$@" >> $out 2>&1 || echo "   (exit $?)" >> $out; }
B="stdbuf -o0 tools/k2bench"
for i in 1 2; do
run LD_PRELOAD=tools/alt/libgas_prev.so $B 16384 512 0.25 16
run X=1 $B 16384 512 0.25 16
done
run LD_PRELOAD=tools/alt/libgas_prev.so $B 16384 512 1.0 16
run X=1 $B 16384 512 1.0 16
run GAS_K2_DEBUG=8 GAS_K2_DUMP=1 $B 16384 512 0.25 16
