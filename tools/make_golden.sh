#!/usr/bin/env bash
# Golden vectors FROM THE REFERENCE ITSELF: runs the compact case list of tests/ref_parity_worker.py through the
# reference's own library (oracle/_ref/libwholegraph_ref.so, built from /root/reference by oracle/build_ref.sh) on a GPU
# box and stores the raw output bytes.  Usage (from the repo root, on the dev container):
#   gpurun -- 'bash tools/make_golden.sh' && cp gpurun_out/reference_gather_scatter_golden.npz tests/golden/
set -euo pipefail
mkdir -p gpurun_out
WG_GOLDEN_SMALL=1 WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so python tests/ref_parity_worker.py gpurun_out/reference_gather_scatter_golden.npz
ls -la gpurun_out/reference_gather_scatter_golden.npz

# The other kernels of the hot path, same idea: the reference's own optimizer kernels (dedup + SGD / LazyAdam / AdaGrad /
# RMSProp through the oracle/_ref hook), its unweighted sampler (on the restated PCG stand-in) and its graph ops, run on the
# seeded case lists of the three workers; tests/test_golden_more.py then pins the oracle to these outputs on CPU.
#   cp gpurun_out/reference_{optimizer,sampler,graph_ops}_golden.npz tests/golden/
for w in optimizer:ref_optimizer_worker sampler:ref_sample_worker graph_ops:ref_graph_ops_worker; do
  name=${w%%:*}; worker=${w##*:}
  WG_GOLDEN_SMALL=1 WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so timeout 900 python tests/$worker.py gpurun_out/reference_${name}_golden.npz \
    && ls -la gpurun_out/reference_${name}_golden.npz || echo "golden for $name: worker failed"
done
