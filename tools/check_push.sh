#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_multi_rank_gpu.py tests/test_reference_binding_gpu.py tests/test_ref_parity_gpu.py -m gpu -q -x -k "gradient or reference or parity" > gpurun_out/push_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/push_pytest.log | cut -c1-300
python tools/bench_ops.py --what adam 2>&1 | tail -2
