#!/bin/bash
# r02s: 2 GPUs exactly as the driver launches it (+ N = 1 beside it), multi-GPU correctness test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02s; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_comm_gpu.py -m gpu -q > $O/pytest_comm.log 2>&1; echo "comm exit $?" >> $O/runs.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 exit $?" >> $O/runs.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 exit $?" >> $O/runs.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 320 --warmup 24 --no-parity > $O/bench_n2_long.json 2> $O/bench_n2_long.err; echo "n2 long exit $?" >> $O/runs.log
