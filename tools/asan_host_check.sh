#!/bin/bash
# Host side of libwholegraph.so under AddressSanitizer + UBSan, on a CPU box (no GPU needed).
# The .cpp files are rebuilt with -fsanitize=address,undefined into a scratch directory, linked with the already built device
# objects, and the CPU-mode C++ caller, the forked 2/3-rank communicator tests and the three randomised differential
# programs (50,000 iterations each, against the reference host code in oracle/_ref) are run against that library.
#   bash tools/asan_host_check.sh            -> prints each program's verdict; any sanitizer report fails the script
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${ASAN_OUT:-$(mktemp -d)}"
CS="$ROOT/wholegraph_b200/csrc"
[ -d "$CS/build" ] || make -C "$CS" -j8 > /dev/null
for f in "$CS"/*.cpp; do
  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -I"$ROOT/include" -I/usr/local/cuda/include -c "$f" -o "$OUT/$(basename "${f%.cpp}").o" &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o "$OUT/libwholegraph.so" "$OUT"/*.o "$CS"/build/*.cu.o \
  -ldl -lpthread -lrt -Xlinker -lasan -Xlinker -lubsan
export ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1
SAN="-fsanitize=address,undefined -g -O1"
g++ -std=c++17 $SAN -I"$ROOT/include" -I/usr/local/cuda/include "$ROOT/tests/cpp/abi_cpp_test.cpp" -o "$OUT/abi_cpp_test" -L"$OUT" -lwholegraph \
  -Wl,-rpath,"$OUT" -L/usr/local/cuda/lib64 -lcudart
"$OUT/abi_cpp_test" cpu 2>&1 | tail -1
declare -A REFSO=([host_diff_test]=ref_host_tensor.so [ops_validation_diff]=ref_host_ops.so [graph_validation_diff]=ref_host_graph.so)
for t in host_diff_test ops_validation_diff graph_validation_diff; do
  [ -f "$ROOT/oracle/_ref/${REFSO[$t]}" ] || { echo "$t: skipped (oracle/_ref/${REFSO[$t]} not built)"; continue; }
  g++ -std=c++17 $SAN -I"$ROOT/include" -I/usr/local/cuda/include "$ROOT/tests/cpp/$t.cpp" -o "$OUT/$t" -ldl
  # LD_LIBRARY_PATH: the reference host code's own NEEDED libwholegraph.so must resolve to THIS sanitized build (one copy in the
  # process: handles are checked against the issuing library's live-object registry)
  LD_LIBRARY_PATH="$OUT" "$OUT/$t" "$OUT/libwholegraph.so" "$ROOT/oracle/_ref/${REFSO[$t]}" 50000 2>&1 | tail -1
done
# the control plane under load (shared-memory mailbox + sockets, intruders) and the stale-handle walk, same sanitized library
g++ -std=c++17 $SAN -I"$ROOT/include" -I/usr/local/cuda/include -I"$CS" "$ROOT/tests/cpp/bootstrap_stress.cpp" -o "$OUT/bootstrap_stress" -L"$OUT" -lwholegraph \
  -Wl,-rpath,"$OUT" -L/usr/local/cuda/lib64 -lcudart
WG_BOOTSTRAP_TIMEOUT_S=120 "$OUT/bootstrap_stress" 4 2000 2>&1 | tail -1
WG_BOOTSTRAP_TIMEOUT_S=120 "$OUT/bootstrap_stress" 3 200 1 2>&1 | tail -1
echo "asan/ubsan host check: clean"
