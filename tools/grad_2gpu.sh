#!/bin/bash
# N-GPU check of the gradient paths: parity tests, then push vs NCCL timing (N = $1, default 2)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_rank_gpu.py -m gpu -q -x -k "gradient" > gpurun_out/grad_${N}gpu_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/grad_${N}gpu_pytest.log | cut -c1-300
for push in 1 0; do
  WG_GRAD_PUSH=$push timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tools/bench_grad_multi.py 2>&1 | grep '^{' 
done
