#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's OWN CPU models of the neighbor samplers -- the host functions its GPU tests
# compare the kernels against (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu) -- for the CPU, from where they lie
# under /root/reference, on top of the restated PCG stand-in (oracle/ref_shim/raft/random/rng_device.cuh) and a four-macro
# gtest stand-in (oracle/ref_shim/gtest/gtest.h; the file only uses EXPECT_EQ / EXPECT_TRUE as argument checks).
# Output: oracle/_ref/ref_host_sampling_model.so (git-ignored), with the extern "C" doors of oracle/ref_sampling_model_hook.cpp.
# tests/test_ref_sampling_model.py pins this repo's oracle (oracle/wm_oracle.c, wm_oracle_weighted.c) against it.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
SRC="$REF/cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu"
OUT="$HERE/_ref/ref_host_sampling_model.so"
[ -f "$SRC" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_sampling_model.sh" ] && [ "$OUT" -nt "$HERE/ref_sampling_model_hook.cpp" ] \
   && [ "$OUT" -nt "$HERE/ref_shim/raft/random/rng_device.cuh" ] && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_sampling_model.so is up to date"; exit 0
fi
g++ -std=c++17 -O1 -fPIC -shared -w -I"$HERE/ref_shim" -I"$REF/cpp/include" -I"$REF/cpp/src" -I"$REF/cpp/tests" -I/usr/local/cuda/include \
  -x c++ "$SRC" "$REF/cpp/src/logger.cpp" "$HERE/ref_sampling_model_hook.cpp" -o "$OUT" \
  -Wl,--unresolved-symbols=ignore-all -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
