#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's HOST replay of the sampler's random streams
# (cpp/src/wholegraph_ops/raft_random_gen.cu: generate_random_positive_int_cpu :27-71 and
# generate_exponential_distribution_negative_float_cpu :73-119 -- plain host code in a .cu file) as C++ for the CPU, from
# where it lies under /root/reference, on top of the restated PCG stand-in (oracle/ref_shim/raft/random/rng_device.cuh).
# The two entry points are renamed ref_* at compile time so they can live next to this repo's own implementations; the
# tensor accessors they call resolve to this repo's libwholegraph.so.  Output: oracle/_ref/ref_host_random.so
# (git-ignored).  tests/test_ref_host_random.py then checks reference code == this repo's library == oracle, on CPU.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
SRC="$REF/cpp/src/wholegraph_ops/raft_random_gen.cu"
OUT="$HERE/_ref/ref_host_random.so"
[ -f "$SRC" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_random.sh" ] && [ "$OUT" -nt "$HERE/ref_shim/raft/random/rng_device.cuh" ] \
   && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_random.so is up to date"; exit 0
fi
g++ -std=c++17 -O2 -fPIC -shared -w -x c++ \
  -Dgenerate_random_positive_int_cpu=ref_generate_random_positive_int_cpu \
  -Dgenerate_exponential_distribution_negative_float_cpu=ref_generate_exponential_distribution_negative_float_cpu \
  -I"$REF/cpp/include" -I"$REF/cpp/src" -I"$HERE/ref_shim" -I/usr/local/cuda/include \
  "$SRC" "$REF/cpp/src/logger.cpp" -o "$OUT" \
  -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
