#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp_summary.txt
for v in G H I J A G; do
  cp tools/alt/libgas_$v.so godot-audio-spatializer_b200/libgas_b200.so
  timeout 120 bash tools/exp_bench.sh pro$v "X=$v" > /dev/null
done
cp tools/alt/libgas_A.so godot-audio-spatializer_b200/libgas_b200.so
