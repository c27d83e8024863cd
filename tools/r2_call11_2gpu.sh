#!/bin/bash
# Round 2, GPU call 11 (2 GPUs): the whole GPU suite in one process on a 2-GPU box (the one-rank-per-GPU cases at world 2 run instead
# of skipping), and the DISTRIBUTED gather three ways: peer-mapped loads (default), forced bucket exchange over NCCL, the reference.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
timeout 2400 python -X faulthandler -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_${R}_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${R}_2gpu.log | cut -c1-300
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 "$@" 2>&1 | grep -E '^\{|rror' | tail -1; }
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['config']['memory_type'], d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
tr bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --shapes c2 --memory-type distributed | tee gpurun_out/bench_2gpu_distributed_ours.json | show "OURS default (peer-mapped)"
WG_FORCE_EXCHANGE=1 tr bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --shapes c2 --memory-type distributed | tee gpurun_out/bench_2gpu_distributed_ours_exchange.json | show "OURS forced NCCL bucket exchange"
tr bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --shapes c2 --memory-type distributed --impl reference | tee gpurun_out/bench_2gpu_distributed_reference.json | show "REFERENCE (its NCCL gather)"
