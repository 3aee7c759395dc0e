#!/bin/bash
# A/B experiments on the streaming mix kernel: per-launch duration from ncu (cold cache, serialised)
# for a few debug configurations.  Usage (on the GPU box): bash tools/k2_experiments.sh
mkdir -p gpurun_out
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --e2e-steps 4"
run() {
  tag=$1; shift
  env "$@" $CMD > gpurun_out/exp_plain_$tag.log 2>&1 || { echo "$tag: plain run failed"; tail -3 gpurun_out/exp_plain_$tag.log; return; }
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mix_stream -c 12 --csv --log-file gpurun_out/exp_$tag.csv $CMD $EXTRA > gpurun_out/exp_ncu_$tag.log 2>&1
  python - <<PY
import csv
lines=[l for l in open("gpurun_out/exp_$tag.csv") if not l.startswith("==")]
v=[float(r["Metric Value"].replace(",","")) for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum"]
print("%-28s n=%d mean=%.2f us min=%.2f us" % ("$tag", len(v), sum(v)/max(1,len(v))/1000, min(v)/1000 if v else 0))
PY
}
run base X=1
run noflush GAS_K2_DEBUG=1
run nofma GAS_K2_DEBUG=2
run noflush_nofma GAS_K2_DEBUG=3
run stages3 GAS_K2_STAGES=3
CMD="$CMD --area-fraction 0"
run area0 X=1
run area0_nofma GAS_K2_DEBUG=2
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --e2e-steps 4 --area-fraction 1"
run area1 X=1
