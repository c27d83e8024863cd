#!/bin/bash
# A/B of the fused optimizer kernel's launch shape + parity of the optimizer paths (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "optim or gradient or embedding or ref_parity or reference_binding" > gpurun_out/opt_pytest.log 2>&1
tail -3 gpurun_out/opt_pytest.log
for u in 1 2; do
  for v in 0 1 2 3 4 5 6 7; do
    echo "== WG_OPT_UNROLL=$u WG_OPT_VARIANT=$v"
    WG_OPT_UNROLL=$u WG_OPT_VARIANT=$v python tools/bench_ops.py --what adam 2>&1 | tail -2 | head -1 | cut -c1-20,100-
  done
done
