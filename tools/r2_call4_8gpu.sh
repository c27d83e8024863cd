#!/bin/bash
# Round 2, GPU call 4 (8 GPUs, charged 8x: keep it short): headline + fp16 shapes both arms, C4 and C5 at size three-way.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
REF=oracle/_ref/libwholegraph_ref.so
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 "$@" 2>&1 | grep -E '^\{|rror' | tail -2; }
show() { python -c "import sys,json; d=json.loads(open('$1').read().strip().split('\n')[-1]); print('$2', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'), [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"; }
echo "=== 1. bench arms at N=$N (C2 weak-scaled + north star + C3)"
tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_${N}gpu_ours_$R.json; show gpurun_out/bench_${N}gpu_ours_$R.json OURS
tr bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --impl reference > gpurun_out/bench_${N}gpu_reference_$R.json; show gpurun_out/bench_${N}gpu_reference_$R.json REF
echo "=== 2. C4: 160M x 512 fp32 DISTRIBUTED LazyAdam, 262144 grads per rank: ours (push), ours (NCCL), reference; uniform and Zipf(1.05)"
tr tools/bench_grad_multi.py --rows-per-gpu 20000000 | cut -c1-600
tr tools/bench_grad_multi.py --rows-per-gpu 20000000 --zipf 1.05 | cut -c1-600
WG_GRAD_PUSH=0 tr tools/bench_grad_multi.py --rows-per-gpu 20000000 | cut -c1-600
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_grad_multi.py --rows-per-gpu 20000000 | cut -c1-600
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_grad_multi.py --rows-per-gpu 20000000 --zipf 1.05 | cut -c1-600
echo "=== 3. C5: 111M nodes / 1B edges, fanout [25,10], 1024 seeds per rank: ours, reference"
tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000 | cut -c1-600
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000 | cut -c1-600
tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000 --seeds 16384 | cut -c1-600
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_sample_multi.py --nodes 111059956 --edges 1000000000 --seeds 16384 | cut -c1-600
