#!/bin/bash
# r03a: the first GPU call after round 2 — everything that was built after round 2's GPU budget ran out, on hardware for the first time:
# the GPU suite (new: filter-tile path of the voice-parallel kernel, resampler, bus graph, resident mix, graph guards), the default
# bench line, the voice-parallel kernel A/B (filter-tile path vs the round-1 per-warp form, GAS_K3_LEGACY=1) on the secondary
# configurations, and one ncu capture of the rebuilt kernel.   gpurun --timeout 1500 -- 'bash tools/run_r03a.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r03a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/runs.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/runs.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; echo "bench exit $?" >> $O/runs.log
GAS_K3_LEGACY=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_k3_legacy.json 2> $O/bench_k3_legacy.err; echo "bench legacy exit $?" >> $O/runs.log
python - <<'PY' > $O/k3_ab.txt 2>&1
import json
def cfgs(p):
    try:
        return {c["config"]: c for c in json.loads(open(p).read().strip().splitlines()[-1])["configs"] or []}
    except Exception as ex:
        return {"error": repr(ex)}
new, old = cfgs("gpurun_out/r03a/bench_driver.json"), cfgs("gpurun_out/r03a/bench_k3_legacy.json")
for k, c in new.items():
    o = old.get(k, {})
    print(f"{c.get('us_per_block')!s:>12} us (filter-tile)  {o.get('us_per_block')!s:>12} us (per-warp)  parity {c.get('parity')}  {k}")
PY
timeout 600 python tools/bench_configs.py > $O/configs_filter_tile.jsonl 2> $O/configs_filter_tile.err; echo "configs exit $?" >> $O/runs.log
GAS_K3_LEGACY=1 timeout 600 python tools/bench_configs.py > $O/configs_per_warp.jsonl 2> $O/configs_per_warp.err; echo "configs legacy exit $?" >> $O/runs.log
# ncu: the voice-parallel kernel on Mode A + filter at the headline size (tools/bench_configs.py, configuration 4); only after the plain run passed
if grep -q "bench exit 0" $O/runs.log; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mix_voice -c 3 -o $O/k3_modeA python tools/bench_configs.py 4 > $O/ncu_k3.log 2>&1; echo "ncu exit $?" >> $O/runs.log
fi
