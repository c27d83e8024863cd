#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_sparse_optimizer_gpu.py tests/test_gather_scatter_gpu.py tests/test_multi_rank_gpu.py -m gpu -q -x -k "optimizer or lazy_adam or arrival or graph or gradient" > gpurun_out/opt_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/opt_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/bench_ops.py --what adam 2>&1 | tail -2
