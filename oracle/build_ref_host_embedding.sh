#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's embedding-layer HOST code (cpp/src/wholememory/embedding.cpp,
# embedding_optimizer.cpp, embedding_cache.cpp) for the CPU, from where it lies under /root/reference, into
# oracle/_ref/ref_host_embedding.so (git-ignored).  Everything it needs from the GPU side stays unresolved at link time
# (--unresolved-symbols=ignore-all) and is bound lazily, so only entry points that never reach the device may be called:
# optimizer creation / parameter names / destruction and cache-policy creation.  tests/test_ref_host_embedding.py
# compares their return codes with this repo's library on the same arguments.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF_ROOT:-/root/reference}"
S="$REF/cpp/src"
OUT="$HERE/_ref/ref_host_embedding.so"
[ -f "$S/wholememory/embedding.cpp" ] || { echo "no reference tree at $REF"; exit 3; }
mkdir -p "$HERE/_ref"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/build_ref_host_embedding.sh" ] && [ "$OUT" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ]; then
  echo "oracle/_ref/ref_host_embedding.so is up to date"; exit 0
fi
g++ -std=c++17 -O1 -fPIC -shared -w -I"$REF/cpp/include" -I"$S" -I"$HERE/ref_shim" -I/usr/local/cuda/include \
  "$S/wholememory/embedding.cpp" "$S/wholememory/embedding_optimizer.cpp" "$S/wholememory/embedding_cache.cpp" "$S/logger.cpp" -o "$OUT" \
  -Wl,-Bsymbolic -Wl,--unresolved-symbols=ignore-all -L"$ROOT/wholegraph_b200/lib" -lwholegraph -Wl,-rpath,'$ORIGIN/../../wholegraph_b200/lib'
echo "built $OUT"
