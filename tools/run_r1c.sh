set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r1c.log
rm -f gpurun_out/exp_summary.txt
bash tools/exp_bench.sh base "X=1"
bash tools/exp_bench.sh pdl "GAS_PDL=1"
bash tools/exp_bench.sh flush0 "GAS_K2_FLUSH=0"
bash tools/exp_bench.sh cost0 "GAS_K2_FIXED_COST=0"
bash tools/exp_bench.sh cost24 "GAS_K2_FIXED_COST=24"
bash tools/exp_bench.sh nofma "GAS_K2_DEBUG=2"
bash tools/exp_bench.sh noflush "GAS_K2_DEBUG=1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-parity --e2e-steps 4 > gpurun_out/ncu_r1c.log 2>&1
cat gpurun_out/exp_summary.txt
