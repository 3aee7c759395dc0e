#!/bin/bash
# Round record, part 2: one `ncu --set full` capture of a kernel (regex $2) after a plain run of the same command.
mkdir -p gpurun_out
tag=${1:-r01}; kern=${2:-k_mix_stream}; name=${3:-k2}
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --e2e-steps 4"
timeout 300 $CMD > gpurun_out/plain2_$tag.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 12 -c 2 -f -o gpurun_out/${name}_$tag $CMD > gpurun_out/ncu_${name}_$tag.log 2>&1
echo "ncu full rc=$?"
