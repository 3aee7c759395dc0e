#!/bin/bash
# A/B runs of bench.py under different schedule knobs (they never change results).  Usage on the GPU box:
#   bash tools/exp_bench.sh TAG "ENV=VAL ENV=VAL" [extra bench args]
# Appends a one-line summary to gpurun_out/exp_summary.txt and keeps the JSON line in gpurun_out/exp_TAG.json
mkdir -p gpurun_out
tag=$1; envs=$2; shift 2
env $envs python bench.py --steps 300 --warmup 20 --no-cpu --no-parity --e2e-steps 4 "$@" > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err || { echo "$tag: FAILED"; tail -5 gpurun_out/exp_$tag.err; exit 1; }
python - "$tag" "$envs" <<'PY' | tee -a gpurun_out/exp_summary.txt
import json,sys
tag,envs=sys.argv[1],sys.argv[2]
d=json.loads(open(f"gpurun_out/exp_{tag}.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("%-14s %-40s step %.2f us  value %.1f G  K2 %.2f us (%.1f%% hbm)  gain %.2f prologue %.2f  K3 %.2f  e2e %.2f ms  clocks %s" % (tag, envs, d["ms_per_step"]*1e3, d["value"]/1e9, r["us_per_launch"], 100*r["frac"], r["other_kernels_us"]["gain_K1"], r["other_kernels_us"]["prologue"], r["other_kernels_us"]["mix_voice_K3"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
