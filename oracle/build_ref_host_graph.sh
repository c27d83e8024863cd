#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's sampling and graph-op entry points
# (cpp/src/wholegraph_ops/{un,}weighted_sample_without_replacement.cpp, cpp/src/graph_ops/append_unique.cpp,
# csr_add_self_loop.cpp) together with its tensor / descriptor code for the CPU, from where they lie under /root/reference,
# into oracle/_ref/ref_host_graph.so (git-ignored).  GPU dispatch targets are replaced by oracle/ref_host_graph_stubs.cpp
# (sentinel 1000).  tests/test_ref_host_graph.py runs tests/cpp/graph_validation_diff.cpp on it and on this repo's library.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
S="$REF/cpp/src"
OUT="$HERE/_ref/ref_host_graph.so"
[ -f "$S/graph_ops/append_unique.cpp" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_graph.sh" ] && [ "$OUT" -nt "$HERE/ref_host_graph_stubs.cpp" ] \
   && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_graph.so is up to date"; exit 0
fi
g++ -std=c++17 -O1 -fPIC -shared -w -I"$REF/cpp/include" -I"$S" -I"$HERE/ref_shim" -I/usr/local/cuda/include \
  "$S/wholegraph_ops/unweighted_sample_without_replacement.cpp" "$S/wholegraph_ops/weighted_sample_without_replacement.cpp" \
  "$S/graph_ops/append_unique.cpp" "$S/graph_ops/csr_add_self_loop.cpp" "$S/wholememory/wholememory_tensor.cpp" \
  "$S/wholememory/tensor_description.cpp" "$S/logger.cpp" "$HERE/ref_host_graph_stubs.cpp" -o "$OUT" \
  -Wl,-Bsymbolic -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
