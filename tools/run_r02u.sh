#!/bin/bash
# r02u: 4 and 8 GPUs exactly as the driver launches them
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02u; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "n$n exit $?" >> $O/runs.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 320 --warmup 24 --no-parity > $O/bench_n8_long.json 2> $O/bench_n8_long.err; echo "n8 long exit $?" >> $O/runs.log
