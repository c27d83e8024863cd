#!/bin/bash
# Multi-GPU call of the next round:   gpurun --gpus N --timeout 1800 -- 'bash tools/round2_multi_gpu.sh'
# Develop at N=2, confirm at N=8 (charged N x the box time).  Everything is device-timed, max over ranks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
R=${ROUND_TAG:-r2}
REF=oracle/_ref/libwholegraph_ref.so
echo "GPUs: $N"
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 "$@" 2>&1 | grep -E '^\{|rror' | tail -2 | cut -c1-600; }

echo "=== 1. one-rank-per-GPU tests (gather/scatter, gradient push + NCCL, sampling, file I/O)"
timeout 1500 python -m pytest tests/test_multi_rank_gpu.py -m gpu -q -x -k "one_rank_per_gpu" > gpurun_out/pytest_${N}gpu_$R.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_${N}gpu_$R.log | cut -c1-300

echo "=== 2. headline gather, N=$N: ours (default), ours with 256-bit accesses, reference kernels"
tr bench.py --gpus $N --steps 20 --warmup 5 | tee -a gpurun_out/bench_${N}gpu_$R.jsonl
WG_VEC32=1 tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e | tee -a gpurun_out/bench_${N}gpu_$R.jsonl
tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --impl reference | tee -a gpurun_out/bench_${N}gpu_$R.jsonl
if [ "$N" -ge 8 ]; then
  echo "--- C3 (1B x 128 fp16) and the north-star shape (1B x 256 fp16) at 8 GPUs"
  tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --dim 128 --dtype fp16 --rows-per-gpu 125000000 | tee -a gpurun_out/bench_${N}gpu_$R.jsonl
  tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --dim 256 --dtype fp16 --rows-per-gpu 125000000 | tee -a gpurun_out/bench_${N}gpu_$R.jsonl
fi

echo "=== 3. gradient apply (C4 write path): peer-store push, NCCL all-to-all, the reference's own pipeline"
ROWS=5000000; [ "$N" -ge 8 ] && ROWS=20000000   # 8 GPUs: the C4-scale variant (160M x 512 fp32 + LazyAdam state = 123 GB per GPU)
tr tools/bench_grad_multi.py --rows-per-gpu $ROWS
WG_GRAD_PUSH=0 tr tools/bench_grad_multi.py --rows-per-gpu $ROWS
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_grad_multi.py --rows-per-gpu $ROWS

echo "=== 4. multi-hop sampling (C5): ours, ours with native env functions, the reference's sampler + append_unique"
SZ=""; [ "$N" -ge 8 ] && SZ="--nodes 111059956 --edges 1000000000"
tr tools/bench_sample_multi.py $SZ
WG_TORCH_NATIVE_ENV=1 tr tools/bench_sample_multi.py $SZ
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_sample_multi.py $SZ
ls -la gpurun_out | tail -6
