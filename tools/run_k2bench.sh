#!/bin/bash
# K2's in-kernel timeline (GAS_K2_DEBUG=8) and block times of the mix side alone, on the GPU box:  bash tools/run_k2bench.sh
# Debug bits: 1 no bus adds, 2 no FMAs, 4 no copies, 8 timeline (GAS_K2_DUMP=1 lists every CTA).
mkdir -p gpurun_out
out=gpurun_out/k2bench.txt; : > $out
run() { echo "## $*" >> $out; timeout 30 env "$@" >> $out 2>&1 || echo "   (exit $?)" >> $out; }
B="stdbuf -o0 tools/k2bench"
run GAS_K2_DEBUG=8 $B 16384 512 0.25 16
run GAS_K2_DEBUG=0 $B 16384 512 0.25 16
run GAS_K2_DEBUG=0 $B 16384 512 0.0 16
run GAS_K2_DEBUG=0 $B 16384 512 1.0 16
run GAS_K2_DEBUG=0 $B 2048 512 0.25 16
cat $out
