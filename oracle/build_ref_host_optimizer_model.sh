#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds oracle/_ref/ref_host_optimizer_model.so: the reference's OWN CPU model of the sparse
# optimizers (class CPUOptimizer in cpp/tests/wholememory_ops/wholememory_embedding_gradient_apply_tests.cu), which lives
# inside a gtest/GPU test file.  The parameter struct and the class are cut out of that file by their first and last
# lines into a TEMPORARY file (removed on exit -- nothing from the reference is kept in the repo), and compiled for the CPU
# together with oracle/ref_optimizer_model_hook.cpp.  -ffp-contract=off like the oracle.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
SRC="$REF/cpp/tests/wholememory_ops/wholememory_embedding_gradient_apply_tests.cu"
OUT="$HERE/_ref/ref_host_optimizer_model.so"
[ -f "$SRC" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_optimizer_model.sh" ] && [ "$OUT" -nt "$HERE/ref_optimizer_model_hook.cpp" ] \
   && [ "$OUT" -nt "$HERE/ref_shim/gtest/gtest.h" ] && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_optimizer_model.so is up to date"; exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
awk '/^struct EmbeddingBackwardTestParams \{/ {on=1} /^void prepare_data_and_reference\(/ {on=0} on' "$SRC" > "$TMP/slice.inc"
grep -q "class CPUOptimizer" "$TMP/slice.inc" || { echo "could not locate CPUOptimizer in $SRC"; exit 4; }
g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -w -DWGREF_OPTIMIZER_MODEL_SLICE="\"$TMP/slice.inc\"" \
  -I"$HERE/ref_shim" -I"$REF/cpp/include" -I/usr/local/cuda/include "$HERE/ref_optimizer_model_hook.cpp" -o "$OUT" \
  -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
