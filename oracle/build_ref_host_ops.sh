#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's op entry points (cpp/src/wholememory_ops/gather_op.cpp, scatter_op.cpp)
# together with its tensor / descriptor code for the CPU, from where they lie under /root/reference, into
# oracle/_ref/ref_host_ops.so (git-ignored).  The GPU functions they dispatch to are replaced by oracle/ref_host_ops_stubs.cpp
# (sentinel return code 1000 = "argument checks passed, call dispatched").  tests/test_ref_host_ops.py runs
# tests/cpp/ops_validation_diff.cpp on it and on this repo's library.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
S="$REF/cpp/src"
OUT="$HERE/_ref/ref_host_ops.so"
[ -f "$S/wholememory_ops/gather_op.cpp" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_ops.sh" ] && [ "$OUT" -nt "$HERE/ref_host_ops_stubs.cpp" ] \
   && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_ops.so is up to date"; exit 0
fi
g++ -std=c++17 -O1 -fPIC -shared -w -I"$REF/cpp/include" -I"$S" -I"$HERE/ref_shim" -I/usr/local/cuda/include \
  "$S/wholememory_ops/gather_op.cpp" "$S/wholememory_ops/scatter_op.cpp" "$S/wholememory/wholememory_tensor.cpp" \
  "$S/wholememory/tensor_description.cpp" "$S/logger.cpp" "$HERE/ref_host_ops_stubs.cpp" -o "$OUT" \
  -Wl,-Bsymbolic -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
