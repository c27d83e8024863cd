#!/usr/bin/env bash
# Golden vectors FROM THE REFERENCE ITSELF: runs the compact case list of tests/ref_parity_worker.py through the
# reference's own library (oracle/_ref/libwholegraph_ref.so, built from /root/reference by oracle/build_ref.sh) on a GPU
# box and stores the raw output bytes.  Usage (from the repo root, on the dev container):
#   gpurun -- 'bash tools/make_golden.sh' && cp gpurun_out/reference_gather_scatter_golden.npz tests/golden/
set -euo pipefail
mkdir -p gpurun_out
WG_GOLDEN_SMALL=1 WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so python tests/ref_parity_worker.py gpurun_out/reference_gather_scatter_golden.npz
ls -la gpurun_out/reference_gather_scatter_golden.npz
