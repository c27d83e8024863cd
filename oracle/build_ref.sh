#!/usr/bin/env bash
# Builds the REFERENCE's own gather/scatter path (rapidsai/wholegraph, unmodified sources where they lie under
# /root/reference) for sm_100 into oracle/_ref/libwholegraph_ref.so -- test infrastructure: the bit-exact GPU
# oracle and the "reference arm" of bench.py.  No reference source is copied into the repo; only objects and the
# .so are written, under the git-ignored oracle/_ref/.  The reference's cmake build needs network (rapids-cmake,
# RAFT via CPM), so the few translation units on the path are compiled directly with a ~70-line RAFT shim
# (oracle/ref_shim) -- recipe from SURVEY.md Appendix B.  The sparse-optimizer kernels
# (functions/embedding_optimizer_func.cu) are built too: that TU includes the embedding-cache header only for
# `CacheLineInfo`, and a declaration-only stand-in for RAFT's warp top-k queue
# (ref_shim/raft/matrix/detail/select_k-inl.cuh) lets it compile; ref_optimizer_hook.cpp exposes
# dedup + optimizer step as one extern "C" test entry.  graph_ops/ (append_unique, csr_add_self_loop) needs only
# integer_utils and is built as is.  The embedding layer (wholememory/embedding*.cpp: create_embedding, optimizers,
# gather_gradient_apply) is built as is too; only the device-cache kernels it can call (embedding_cache_func.cu,
# gather_cached_func.cu: real RAFT select_k) are replaced by loud NOT_IMPLEMENTED stubs (ref_cache_stubs.cpp).
# The UNWEIGHTED sampler (wholegraph_ops/unweighted_*, raft_random_gen.cu) is built as is on top of a RESTATED PCG generator
# (ref_shim/raft/random/rng_device.cuh -- same stream as oracle/wm_oracle.c, not verified against RAFT): the reference's
# selection kernels are real, the random stream is the restated one.  The weighted sampler needs RAFT's select_k: NOT built.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
[ -d "$REF/cpp/src" ] || { echo "no reference tree at $REF"; exit 3; }
if [ -f "$OUT/libwholegraph_ref.so" ] && [ "$OUT/libwholegraph_ref.so" -nt "$HERE/build_ref.sh" ] && [ "$OUT/libwholegraph_ref.so" -nt "$HERE/ref_stubs.cpp" ] \
   && [ "$OUT/libwholegraph_ref.so" -nt "$HERE/ref_optimizer_hook.cpp" ] && [ "$OUT/libwholegraph_ref.so" -nt "$HERE/ref_cache_stubs.cpp" ]; then
  echo "oracle/_ref/libwholegraph_ref.so is up to date"; exit 0
fi
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
INC="-I$REF/cpp/include -I$REF/cpp/src -I$HERE/ref_shim -I/usr/include -I/usr/local/cuda/include"
NVFLAGS="-std=c++17 -O3 -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -w $INC"
CXXFLAGS="-std=c++17 -O2 -fPIC -w $INC"
S="$REF/cpp/src"
CPP="cuda_macros.cpp logger.cpp
 wholememory/communicator.cpp wholememory/nccl_comms.cpp wholememory/memory_handle.cpp wholememory/wholememory.cpp
 wholememory/wholememory_tensor.cpp wholememory/tensor_description.cpp wholememory/env_func_ptrs.cpp
 wholememory/initialize.cpp wholememory/system_info.cpp wholememory/global_reference.cpp
 wholememory_ops/gather_op.cpp wholememory_ops/scatter_op.cpp wholememory_ops/thrust_allocator.cpp
 graph_ops/append_unique.cpp graph_ops/csr_add_self_loop.cpp
 wholememory/embedding.cpp wholememory/embedding_optimizer.cpp wholememory/embedding_cache.cpp
 wholegraph_ops/unweighted_sample_without_replacement.cpp"
CU="wholememory_ops/gather_op_impl_mapped.cu wholememory_ops/gather_op_impl_nccl.cu
 wholememory_ops/scatter_op_impl_mapped.cu wholememory_ops/scatter_op_impl_nccl.cu
 wholememory_ops/functions/gather_func.cu wholememory_ops/functions/scatter_func.cu
 wholememory_ops/functions/bucket_ids_func.cu wholememory_ops/functions/exchange_ids_nccl_func.cu
 wholememory_ops/functions/exchange_embeddings_nccl_func.cu wholememory_ops/functions/sort_indices_func.cu
 wholememory_ops/functions/embedding_optimizer_func.cu wholememory_ops/functions/map_indices_func.cu
 graph_ops/append_unique_impl.cu graph_ops/csr_add_self_loop_impl.cu
 wholegraph_ops/unweighted_sample_without_replacement_impl_mapped.cu
 wholegraph_ops/unweighted_sample_without_replacement_impl_nccl.cu wholegraph_ops/raft_random_gen.cu
 wholememory_ops/functions/gather_func_impl_floating_data_int32_indices.cu
 wholememory_ops/functions/gather_func_impl_floating_data_int64_indices.cu
 wholememory_ops/functions/gather_func_impl_integer_data_int32_indices.cu
 wholememory_ops/functions/gather_func_impl_integer_data_int64_indices.cu
 wholememory_ops/functions/scatter_func_impl_floating_data_int32_indices.cu
 wholememory_ops/functions/scatter_func_impl_floating_data_int64_indices.cu
 wholememory_ops/functions/scatter_func_impl_integer_data_int32_indices.cu
 wholememory_ops/functions/scatter_func_impl_integer_data_int64_indices.cu"
MK="$OBJ/Makefile"
{
  echo "all: objs"
  OBJS=""
  for f in $CPP; do o="$OBJ/$(echo "$f" | tr '/' '_').o"; OBJS="$OBJS $o"; printf '%s: %s\n\tg++ %s -c $< -o $@\n' "$o" "$S/$f" "$CXXFLAGS"; done
  for f in $CU; do o="$OBJ/$(echo "$f" | tr '/' '_').o"; OBJS="$OBJS $o"; printf '%s: %s\n\t%s %s -c $< -o $@\n' "$o" "$S/$f" "$NVCC" "$NVFLAGS"; done
  o="$OBJ/ref_stubs.o"; OBJS="$OBJS $o"; printf '%s: %s\n\tg++ %s -c $< -o $@\n' "$o" "$HERE/ref_stubs.cpp" "$CXXFLAGS"
  o="$OBJ/ref_optimizer_hook.o"; OBJS="$OBJS $o"; printf '%s: %s\n\tg++ %s -c $< -o $@\n' "$o" "$HERE/ref_optimizer_hook.cpp" "$CXXFLAGS"
  o="$OBJ/ref_cache_stubs.o"; OBJS="$OBJS $o"; printf '%s: %s\n\tg++ %s -c $< -o $@\n' "$o" "$HERE/ref_cache_stubs.cpp" "$CXXFLAGS"
  echo "objs:$OBJS"
  echo "OBJS=$OBJS"
} > "$MK"
make -f "$MK" -j"$(nproc)" objs
OBJS=$(grep '^OBJS=' "$MK" | cut -d= -f2-)
# libnvidia-ml: the reference probes GPU fabric info through NVML at communicator creation (system_info.cpp:139-154)
NVML="-lnvidia-ml"
[ -e /usr/lib/x86_64-linux-gnu/libnvidia-ml.so.1 ] || NVML="-L/usr/local/cuda/lib64/stubs -lnvidia-ml"
$NVCC -shared -gencode arch=compute_100,code=sm_100 -o "$OUT/libwholegraph_ref.so" $OBJS -lcuda -lnccl $NVML -L/usr/local/cuda/lib64/stubs -ldl -lpthread
echo "built $OUT/libwholegraph_ref.so"
