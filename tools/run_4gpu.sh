#!/bin/bash
# 4-GPU measurements: gradient apply push vs NCCL, then the headline gather bench (default kernel)
cd "$(dirname "$0")/.."
N=${1:-4}
mkdir -p gpurun_out
for push in 1 0; do
  WG_GRAD_PUSH=$push timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    tools/bench_grad_multi.py 2>&1 | grep '^{'
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_${N}gpu.err | tee gpurun_out/bench_${N}gpu.json | cut -c1-300
