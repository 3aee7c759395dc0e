#!/bin/bash
# BASELINE.json configs[4]: stress sweep on one GPU (voices x frames), one bench.py line each (no CPU arm, no parity pass)
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for V in 256 4096 16384 65536; do for F in 128 512 2048; do
  timeout 150 python bench.py --voices $V --frames $F --steps 300 --warmup 20 --no-cpu --no-parity --e2e-steps 4 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl || echo "{\"failed\": [$V, $F]}" >> gpurun_out/sweep.jsonl
done; done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    d=json.loads(l)
    if "failed" in d: print("FAILED", d); continue
    c=d["config"]; r=d["roofline"]
    print(f'{c["voices_per_gpu"]:6d} x {c["frames"]:4d}  step {d["ms_per_step"]*1e3:8.2f} us  {d["value"]/1e9:7.1f} G vf/s  K2 {r["us_per_launch"]:7.2f} us {100*r["frac"]:5.1f}% of peak  step/HBM {100*r["step_frac_of_hbm_peak"]:5.1f}%')
PY
