#!/bin/bash
# Round 2, GPU call 6 (1 GPU): hot-row (long-run) optimizer kernel and the rotated gradient push -- correctness, then timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REF=oracle/_ref/libwholegraph_ref.so
timeout 1200 python -m pytest tests/test_sparse_optimizer_gpu.py tests/test_zz_ref_optimizer_parity_gpu.py tests/test_zz_training_autograd_gpu.py "tests/test_multi_rank_gpu.py::test_single_rank" "tests/test_multi_rank_gpu.py::test_ranks_sharing_one_gpu_mapped_memory" -m gpu -q -p no:cacheprovider -k "not sampling and not file_io" > gpurun_out/pytest_call6.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_call6.log | cut -c1-500
timeout 600 python tools/bench_ops.py --what adam 2>&1 | tail -3
WHOLEGRAPH_B200_LIB=$REF timeout 600 python tools/bench_ops.py --what adam 2>&1 | tail -3
timeout 600 python tools/bench_ops.py --what sample 2>&1 | head -2
WG_TORCH_NATIVE_ENV=1 timeout 600 python tools/bench_ops.py --what sample 2>&1 | head -2
