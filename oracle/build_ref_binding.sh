#!/usr/bin/env bash
# Drop-in proof (test infrastructure): cythonize the reference's UNMODIFIED binding
# (/root/reference/python/pylibwholegraph/pylibwholegraph/binding/wholememory_binding.pyx) against THIS repo's headers and
# link it to THIS repo's libwholegraph.so.  Output: oracle/_ref/refbinding/wholememory_binding*.so (git-ignored).
# tests/test_reference_binding_gpu.py then drives our kernels through the reference's own cython module.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
PYX="${REF_ROOT:-/root/reference}/python/pylibwholegraph/pylibwholegraph/binding/wholememory_binding.pyx"
OUT="$HERE/_ref/refbinding"
[ -f "$PYX" ] || { echo "no reference binding at $PYX"; exit 3; }
mkdir -p "$OUT"
SO="$OUT/wholememory_binding$(python3-config --extension-suffix)"
if [ -f "$SO" ] && [ "$SO" -nt "$ROOT/wholegraph_b200/lib/libwholegraph.so" ] && [ "$SO" -nt "$HERE/build_ref_binding.sh" ]; then
  echo "reference binding is up to date"; exit 0
fi
cython --cplus -3 "$PYX" -o "$OUT/wholememory_binding.cpp"
g++ -std=c++17 -O1 -fPIC -shared -w "$OUT/wholememory_binding.cpp" -o "$SO" $(python3-config --includes) \
    -I"$ROOT/include" -I/usr/local/cuda/include -L"$ROOT/wholegraph_b200/lib" -lwholegraph \
    -Wl,-rpath,'$ORIGIN/../../../wholegraph_b200/lib'
rm -f "$OUT/wholememory_binding.cpp"
echo "built $SO"
