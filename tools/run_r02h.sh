#!/bin/bash
# r02h: ncu --set full of the four kernels of the step (one capture command, after the plain run exited 0) + K2 timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02h; mkdir -p $O
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --no-configs --e2e-steps 2"
timeout 300 $CMD > $O/plain.log 2>&1; echo "plain exit $?" >> $O/runs.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gain|k_prologue|k_mix_stream|k_mix_voice' -s 24 -c 8 -f -o $O/step_full $CMD > $O/ncu_full.log 2>&1; echo "ncu exit $?" >> $O/runs.log
make -C tools -s k2bench && bash tools/run_k2bench.sh > /dev/null 2>&1; cp gpurun_out/k2bench.txt $O/ 2>/dev/null
