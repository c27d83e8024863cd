#!/bin/bash
# Builds the native torch env-function module in-tree:
#   wholegraph_b200/lib/wholegraph_b200_torch_ext<EXT_SUFFIX>
# Plain g++ against the installed torch's headers/libs (no JIT cache: the .so must travel with the repo snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../../.." && pwd)"
OUT_DIR="$ROOT/wholegraph_b200/lib"
mkdir -p "$OUT_DIR"
PY="${PYTHON:-python}"
read -r PY_INC EXT_SUFFIX TORCH_DIR CXX11_ABI < <("$PY" - <<'EOF'
import os, sysconfig, torch
print(sysconfig.get_paths()["include"], sysconfig.get_config_var("EXT_SUFFIX"), os.path.dirname(torch.__file__),
      int(torch._C._GLIBCXX_USE_CXX11_ABI))
EOF
)
OUT="$OUT_DIR/wholegraph_b200_torch_ext$EXT_SUFFIX"
SRC="$HERE/torch_env.cpp"
if [ -f "$OUT" ] && [ "$OUT" -nt "$SRC" ] && [ "$OUT" -nt "$ROOT/include/wholememory/env_func_ptrs.h" ]; then
  exit 0
fi
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
g++ -std=c++17 -O2 -fPIC -shared -w \
  -DTORCH_EXTENSION_NAME=wholegraph_b200_torch_ext -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI="$CXX11_ABI" \
  -I"$ROOT/include" -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PY_INC" -I"$CUDA_INC" \
  "$SRC" -o "$OUT" \
  -L"$TORCH_DIR/lib" -Wl,-rpath,"$TORCH_DIR/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch -ltorch_python
echo "built $OUT"
