#!/bin/bash
# r02k: ncu --set full (with source) of the step kernel with control work, and of the stand-alone planner
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02k; mkdir -p $O
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --no-configs --e2e-steps 2"
timeout 300 $CMD > $O/plain.log 2>&1; echo "plain exit $?" >> $O/runs.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step|k_plan' -s 20 -c 3 -f -o $O/step_full $CMD > $O/ncu_full.log 2>&1; echo "ncu exit $?" >> $O/runs.log
