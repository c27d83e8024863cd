#!/bin/bash
# Round 2, GPU call 8 (8 GPUs, short): C4 after the owner-rotation and hot-row changes, C5 with the native env default,
# and the one-rank-per-GPU gradient test at 8 ranks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REF=oracle/_ref/libwholegraph_ref.so
N=$(nvidia-smi -L | wc -l)
tr() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 "$@" 2>&1 | grep -E '^\{|rror' | tail -2 | cut -c1-520; }
echo "=== C4 at size: ours push uniform / zipf, ours NCCL uniform"
tr tools/bench_grad_multi.py --rows-per-gpu 20000000
tr tools/bench_grad_multi.py --rows-per-gpu 20000000 --zipf 1.05
WG_GRAD_PUSH=0 tr tools/bench_grad_multi.py --rows-per-gpu 20000000
echo "=== C5 at size: ours (native env default), ours with Python callbacks, 1024 and 16384 seeds"
tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000
WG_TORCH_NATIVE_ENV=0 tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000
tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000 --seeds 16384
echo "=== gradient scenario, one rank per GPU, 8 ranks (push) + 4 ranks (NCCL)"
timeout 600 python -m pytest "tests/test_multi_rank_gpu.py::test_one_rank_per_gpu" "tests/test_multi_rank_gpu.py::test_one_rank_per_gpu_gradient_over_nccl" -m gpu -q -p no:cacheprovider -k "gradient" 2>&1 | tail -3
