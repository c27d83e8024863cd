#!/bin/bash
# Round 2, GPU call 3 (2 GPUs): what caps peer reads?  One process drives both GPUs (tools/rowmove_lab.cu), so reads can be
# one-directional (GPU0 <- GPU1, GPU1 idle) or bidirectional (both at once), gathers (peer loads) or scatters (peer
# stores), with NVML NVLink data/raw counters around every timed loop; then the real 2-rank tests and bench arms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
REF=oracle/_ref/libwholegraph_ref.so
L=wholegraph_b200/lib/rowmove_lab
echo "=== 0. tests that failed in call 2 (test bugs / library load order), optimizer golden, reference arm at N=1"
timeout 900 python -m pytest tests/test_access_width_gpu.py tests/test_ref_parity_gpu.py tests/test_zz_ref_graph_ops_parity_gpu.py tests/test_zz_ref_optimizer_parity_gpu.py tests/test_zz_ref_sampling_parity_gpu.py tests/test_zz_reference_binding_full_gpu.py tests/test_reference_binding_gpu.py tests/test_device_reference_gpu.py tests/test_embedding_gather_gpu.py tests/test_stale_handles.py "tests/test_multi_rank_gpu.py::test_rank_local_failures_reach_every_rank" -m gpu -q -p no:cacheprovider > gpurun_out/pytest_retry_$R.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_retry_$R.log | cut -c1-400
WG_GOLDEN_SMALL=1 WHOLEGRAPH_B200_LIB=$REF timeout 900 python tests/ref_optimizer_worker.py gpurun_out/reference_optimizer_golden.npz 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference_$R.err; tail -1 gpurun_out/bench_reference_$R.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('REF N=1', d['value'], d['ms_per_step'], d['e2e']['value'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"
timeout 300 python tools/bench_ops.py --what sample 2>&1 | head -2
nvidia-smi topo -m 2>&1 | head -8
nvidia-smi nvlink -s -i 0 2>&1 | head -8
echo "=== 1. peer reads, 1 KiB rows: one direction vs both directions, launch-shape / access-width / policy variants"
timeout 300 $L --row-bytes 1024 --rows 20000000 --mode uni --set link | tee gpurun_out/lab_uni_rb1024.txt
timeout 300 $L --row-bytes 1024 --rows 20000000 --mode bidir --set link | tee gpurun_out/lab_bidir_rb1024.txt
echo "=== 2. sequential rows (contiguous 1 GiB read over the link) and bigger batches"
timeout 120 $L --row-bytes 1024 --rows 20000000 --mode uni --set default --pattern seq
timeout 120 $L --row-bytes 1024 --rows 20000000 --mode bidir --set default --pattern seq
timeout 120 $L --row-bytes 1024 --rows 20000000 --mode bidir --set default --n 4194304
timeout 120 $L --row-bytes 4096 --rows 5000000 --mode bidir --set default
echo "=== 3. smaller rows over the link"
timeout 200 $L --row-bytes 512 --rows 40000000 --mode bidir --set link | tee gpurun_out/lab_bidir_rb512.txt
timeout 200 $L --row-bytes 256 --rows 80000000 --mode bidir --set link | tee gpurun_out/lab_bidir_rb256.txt
echo "=== 4. peer STORES (scatter), one direction and both"
timeout 200 $L --row-bytes 1024 --rows 20000000 --mode uni --op scatter --set link | tee gpurun_out/lab_uni_scatter_rb1024.txt
timeout 200 $L --row-bytes 1024 --rows 20000000 --mode bidir --op scatter --set link | tee gpurun_out/lab_bidir_scatter_rb1024.txt
echo "=== 5. ncu NVLink counters on the lab (kernels serialised by the profiler: byte counts, not rates)"
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" | head -40 > gpurun_out/ncu_nvlink_metric_names.txt; wc -l gpurun_out/ncu_nvlink_metric_names.txt
M=$(awk '{print $1}' gpurun_out/ncu_nvlink_metric_names.txt | grep -E "bytes" | head -24 | paste -sd, -)
[ -n "$M" ] && timeout 300 ncu --metrics gpu__time_duration.sum,$M --clock-control none -k regex:row_move_vec -s 6 -c 2 --csv --log-file gpurun_out/${R}_nvlink_uni_rb1024.csv $L --row-bytes 1024 --rows 20000000 --mode uni --set default > /dev/null 2>&1
head -c 3000 gpurun_out/${R}_nvlink_uni_rb1024.csv | tail -c 2200
echo "=== 6. one-rank-per-GPU tests (real NCCL all-to-all, peer push, sampling, file I/O)"
timeout 1200 python -m pytest tests/test_multi_rank_gpu.py -m gpu -q -p no:cacheprovider -k "one_rank_per_gpu" > gpurun_out/pytest_2gpu_$R.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_2gpu_$R.log | cut -c1-300
echo "=== 7. bench arms at N=2"
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 "$@" 2>&1 | grep -E '^\{|rror' | tail -2; }
tr bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_ours_$R.json; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_2gpu_ours_$R.json').read().strip().split('\n')[-1]); print('OURS', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"
tr bench.py --gpus 2 --steps 20 --warmup 5 --impl reference --no-e2e > gpurun_out/bench_2gpu_reference_$R.json; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_2gpu_reference_$R.json').read().strip().split('\n')[-1]); print('REF', d['value'], d['ms_per_step'], [(s['key'], s['value'], s['ms_per_step'], s['frac']) for s in d['shapes']])"
echo "=== 8. gradient apply at N=2: peer push, NCCL, the reference pipeline"
tr tools/bench_grad_multi.py --rows-per-gpu 5000000 | cut -c1-500
WG_GRAD_PUSH=0 tr tools/bench_grad_multi.py --rows-per-gpu 5000000 | cut -c1-500
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_grad_multi.py --rows-per-gpu 5000000 | cut -c1-500
echo "=== 9. multi-hop sampling at N=2: ours, the reference"
tr tools/bench_sample_multi.py | cut -c1-500
WHOLEGRAPH_B200_LIB=$REF tr tools/bench_sample_multi.py | cut -c1-500
