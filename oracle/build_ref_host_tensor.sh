#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's HOST-side descriptor and tensor-view code
# (cpp/src/wholememory/tensor_description.cpp, cpp/src/wholememory/wholememory_tensor.cpp) for the CPU, from where it lies
# under /root/reference, into oracle/_ref/ref_host_tensor.so (git-ignored).  -Bsymbolic keeps its internal calls inside
# the object; the WholeMemory-handle functions it references (only reached for handle-backed tensors, which the test does
# not create) resolve to this repo's libwholegraph.so.  tests/test_ref_host_tensor.py runs tests/cpp/host_diff_test.cpp
# on both libraries: same randomised inputs, every return value / descriptor / pointer compared.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
S="$REF/cpp/src"
OUT="$HERE/_ref/ref_host_tensor.so"
[ -f "$S/wholememory/tensor_description.cpp" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_tensor.sh" ] && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_tensor.so is up to date"; exit 0
fi
g++ -std=c++17 -O2 -fPIC -shared -w -I"$REF/cpp/include" -I"$S" -I"$HERE/ref_shim" -I/usr/local/cuda/include \
  "$S/wholememory/tensor_description.cpp" "$S/wholememory/wholememory_tensor.cpp" "$S/logger.cpp" -o "$OUT" \
  -Wl,-Bsymbolic -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
