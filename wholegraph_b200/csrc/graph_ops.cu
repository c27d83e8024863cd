/*
 * Graph helper ops that sit between sampling hops: graph_append_unique and csr_add_self_loop.
 * Replaces reference cpp/src/graph_ops/append_unique_func.cuh:47-353 and csr_add_self_loop_func.cuh:24-58.
 *
 * append_unique(targets, neighbors) -> unique = targets ++ (distinct neighbors not among targets),
 *                                      mapping[i] = position of neighbors[i] in unique.
 * The reference orders the appended neighbors by hash-table slot (arbitrary; its tests compare sorted sets,
 * python/.../tests/wholegraph_torch/ops/test_graph_append_unique.py).  Here they are appended in order of FIRST
 * OCCURRENCE, which is deterministic: an open-addressing table keeps, per key, the minimum of (target index) and
 * (T + first neighbor position); a neighbor is "new" iff that minimum is its own T + position; an exclusive scan of the
 * new-flags gives the final ids.  One host sync (the output size), as the allocation callback ABI requires.
 */
#include "ops_internal.hpp"
#include "wm_internal.hpp"

#include <cub/device/device_scan.cuh>

namespace wm {
namespace {

template <typename KeyT>
__device__ __forceinline__ uint32_t hash_key(KeyT k, uint32_t mask)
{
  uint64_t x = (uint64_t)k;
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x & mask;
}

template <typename KeyT>
__device__ __forceinline__ KeyT cas_key(KeyT* p, KeyT cmp, KeyT val);
template <>
__device__ __forceinline__ int32_t cas_key<int32_t>(int32_t* p, int32_t cmp, int32_t val)
{
  return atomicCAS(p, cmp, val);
}
template <>
__device__ __forceinline__ int64_t cas_key<int64_t>(int64_t* p, int64_t cmp, int64_t val)
{
  return (int64_t)atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)val);
}

constexpr int kEmptyVal = 0x7fffffff;

/* insert keys[i] with candidate value base + i (minimum wins); returns nothing */
template <typename KeyT>
__global__ void insert_kernel(const KeyT* __restrict__ keys, int n, int base, KeyT* table_keys, int* table_vals, uint32_t mask)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KeyT k  = keys[i];
  uint32_t slot = hash_key(k, mask);
  for (;;) {
    KeyT cur = table_keys[slot];
    if (cur == (KeyT)-1) cur = cas_key<KeyT>(table_keys + slot, (KeyT)-1, k) == (KeyT)-1 ? k : table_keys[slot];
    if (cur == k) {
      atomicMin(table_vals + slot, base + i);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

template <typename KeyT>
__device__ __forceinline__ uint32_t find_slot(KeyT k, const KeyT* table_keys, uint32_t mask)
{
  uint32_t slot = hash_key(k, mask);
  while (table_keys[slot] != k) slot = (slot + 1) & mask;
  return slot;
}

template <typename KeyT>
__global__ void flag_new_kernel(const KeyT* __restrict__ neighbors, int n, int T, const KeyT* table_keys, const int* table_vals,
                                uint32_t mask, int* flags)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  flags[i] = (i < n && table_vals[find_slot(neighbors[i], table_keys, mask)] == T + i) ? 1 : 0; /* flags[n] = 0: scan tail = total */
}

template <typename KeyT>
__global__ void emit_unique_kernel(const KeyT* __restrict__ neighbors, int n, int T, const KeyT* table_keys, int* table_vals,
                                   uint32_t mask, const int* __restrict__ flags, const int* __restrict__ pos, KeyT* unique_out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || flags[i] == 0) return;
  const KeyT k      = neighbors[i];
  const int id      = T + pos[i];
  unique_out[id]    = k;
  table_vals[find_slot(k, table_keys, mask)] = id; /* final id replaces "T + first position" */
}

template <typename KeyT>
__global__ void mapping_kernel(const KeyT* __restrict__ neighbors, int n, const KeyT* table_keys, const int* table_vals, uint32_t mask,
                               int* mapping)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mapping[i] = table_vals[find_slot(neighbors[i], table_keys, mask)];
}

__global__ void fill_int_kernel(int* p, int v, size_t n)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

template <typename KeyT>
void append_unique_typed(const void* targets, int T, const void* neighbors, int N, void* unique_ctx, int* mapping, wholememory_dtype_t dt,
                         wholememory_env_func_t* env, cudaStream_t s)
{
  const KeyT* tg = static_cast<const KeyT*>(targets);
  const KeyT* nb = static_cast<const KeyT*>(neighbors);
  size_t slots   = 64;
  while (slots < (size_t)(T + N) * 2) slots <<= 1;
  const uint32_t mask = (uint32_t)(slots - 1);
  temp_buffer keys_b(env), vals_b(env), flags_b(env), pos_b(env), cub_b(env);
  KeyT* tk = static_cast<KeyT*>(keys_b.device(slots, dt));
  int* tv  = static_cast<int*>(vals_b.device(slots, WHOLEMEMORY_DT_INT));
  WM_CUDA(cudaMemsetAsync(tk, 0xff, slots * sizeof(KeyT), s));
  fill_int_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(tv, kEmptyVal, slots);
  if (T > 0) insert_kernel<KeyT><<<(T + 255) / 256, 256, 0, s>>>(tg, T, 0, tk, tv, mask);
  if (N > 0) insert_kernel<KeyT><<<(N + 255) / 256, 256, 0, s>>>(nb, N, T, tk, tv, mask);
  int* flags = static_cast<int*>(flags_b.device((size_t)N + 1, WHOLEMEMORY_DT_INT));
  int* pos   = static_cast<int*>(pos_b.device((size_t)N + 1, WHOLEMEMORY_DT_INT));
  flag_new_kernel<KeyT><<<(N + 1 + 255) / 256, 256, 0, s>>>(nb, N, T, tk, tv, mask, flags);
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, flags, pos, N + 1, s);
  void* cub_tmp = cub_b.device(cub_bytes, WHOLEMEMORY_DT_INT8);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flags, pos, N + 1, s);
  int fresh = 0;
  read_back_sync(&fresh, pos + N, sizeof(int), s);
  KeyT* uniq = static_cast<KeyT*>(output_alloc(env, unique_ctx, (size_t)T + fresh, dt));
  if (T > 0) WM_CUDA(cudaMemcpyAsync(uniq, tg, (size_t)T * sizeof(KeyT), cudaMemcpyDeviceToDevice, s));
  if (N > 0) {
    emit_unique_kernel<KeyT><<<(N + 255) / 256, 256, 0, s>>>(nb, N, T, tk, tv, mask, flags, pos, uniq);
    if (mapping != nullptr) mapping_kernel<KeyT><<<(N + 255) / 256, 256, 0, s>>>(nb, N, tk, tv, mask, mapping);
  }
  WM_CUDA(cudaGetLastError());
  WM_CUDA(cudaStreamSynchronize(s)); /* temporaries are released on return */
}

/* one CTA per row: row i keeps its neighbours and gains the edge (i, i) in front (reference kernel :24-45) */
__global__ void add_self_loop_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col, int* out_row_ptr, int* out_col)
{
  const int row = blockIdx.x;
  const int b = row_ptr[row], e = row_ptr[row + 1];
  if (threadIdx.x == 0) {
    out_row_ptr[row] = b + row;
    if (row == (int)gridDim.x - 1) out_row_ptr[row + 1] = e + row + 1;
    out_col[b + row] = row;
  }
  for (int k = threadIdx.x; k < e - b; k += blockDim.x) out_col[b + row + 1 + k] = col[b + k];
}

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t graph_append_unique(wholememory_tensor_t target_nodes_tensor,
                                             wholememory_tensor_t neighbor_nodes_tensor,
                                             void* output_unique_node_memory_context,
                                             wholememory_tensor_t output_neighbor_raw_to_unique_mapping_tensor,
                                             wholememory_env_func_t* p_env_fns,
                                             void* stream)
{
  return wm::guarded("graph_append_unique", [&]() -> wholememory_error_code_t {
    using namespace wm;
    WM_REQUIRE_LIVE(target_nodes_tensor);
    WM_REQUIRE_LIVE(neighbor_nodes_tensor);
    if (output_neighbor_raw_to_unique_mapping_tensor != nullptr) WM_REQUIRE_LIVE(output_neighbor_raw_to_unique_mapping_tensor);
    /* argument checks in the reference's order and with its codes (graph_ops/append_unique.cpp:28-72) ... */
    if (!is_1d(target_nodes_tensor)) {
      WM_ERROR("target_nodes_tensor should be 1D tensor.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (!is_1d(neighbor_nodes_tensor)) {
      WM_ERROR("neighbor_nodes_tensor should be 1D tensor.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    int* mapping = nullptr;
    bool want_mapping = false;
    if (output_neighbor_raw_to_unique_mapping_tensor != nullptr) { /* the reference requires a tensor object; "None" is a 0-dim one */
      auto md = *wholememory_tensor_get_tensor_description(output_neighbor_raw_to_unique_mapping_tensor);
      if (md.dim != 1 && md.dim != 0) {
        WM_ERROR("output_neighbor_raw_to_unique_mapping_tensor should be 1D tensor or None.");
        return WHOLEMEMORY_INVALID_INPUT;
      }
      if (md.dim == 1 && md.dtype != WHOLEMEMORY_DT_INT) {
        WM_ERROR("output_neighbor_raw_to_unique_mapping_tensor should be int tensor or None.");
        return WHOLEMEMORY_INVALID_INPUT;
      }
      want_mapping = md.dim == 1;
    }
    if (!views_as_array(target_nodes_tensor) || !views_as_array(neighbor_nodes_tensor)) {
      WM_ERROR("Input target_nodes_tensor / neighbor_nodes_tensor convert to array failed.");
      return WHOLEMEMORY_LOGIC_ERROR;
    }
    auto td = *wholememory_tensor_get_tensor_description(target_nodes_tensor);
    auto nd = *wholememory_tensor_get_tensor_description(neighbor_nodes_tensor);
    if (td.dtype != nd.dtype) {
      WM_ERROR("target_nodes_dtype should be the same with neighbor_nodes_dtype");
      return WHOLEMEMORY_LOGIC_ERROR;
    }
    /* ... then the point where the reference dispatches to its GPU translation unit, whose failures are all LOGIC_ERROR
     * (append_unique_impl.cu:36-56) */
    require_cuda("graph_append_unique");
    WM_EXPECT(output_unique_node_memory_context != nullptr && p_env_fns != nullptr, WHOLEMEMORY_INVALID_INPUT,
              "graph_append_unique needs an output memory context and env functions");
    WM_EXPECT(td.dtype == WHOLEMEMORY_DT_INT || td.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR,
              "target / neighbor nodes must be int32 or int64");
    if (want_mapping) {
      auto md = *wholememory_tensor_get_tensor_description(output_neighbor_raw_to_unique_mapping_tensor);
      WM_EXPECT(md.sizes[0] == nd.sizes[0], WHOLEMEMORY_INVALID_INPUT,
                "output_neighbor_raw_to_unique_mapping_tensor should have one entry per neighbor");
      mapping = static_cast<int*>(wholememory_tensor_get_data_pointer(output_neighbor_raw_to_unique_mapping_tensor));
    }
    WM_EXPECT(td.sizes[0] + nd.sizes[0] < ((int64_t)1 << 30), WHOLEMEMORY_INVALID_VALUE, "too many nodes for append_unique");
    const void* tg = wholememory_tensor_get_data_pointer(target_nodes_tensor);
    const void* nb = wholememory_tensor_get_data_pointer(neighbor_nodes_tensor);
    auto s         = static_cast<cudaStream_t>(stream);
    if (td.dtype == WHOLEMEMORY_DT_INT64)
      append_unique_typed<int64_t>(tg, (int)td.sizes[0], nb, (int)nd.sizes[0], output_unique_node_memory_context, mapping, td.dtype, p_env_fns, s);
    else
      append_unique_typed<int32_t>(tg, (int)td.sizes[0], nb, (int)nd.sizes[0], output_unique_node_memory_context, mapping, td.dtype, p_env_fns, s);
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t csr_add_self_loop(wholememory_tensor_t csr_row_ptr_tensor,
                                           wholememory_tensor_t csr_col_ptr_tensor,
                                           wholememory_tensor_t output_csr_row_ptr_tensor,
                                           wholememory_tensor_t output_csr_col_ptr_tensor,
                                           void* stream)
{
  return wm::guarded("csr_add_self_loop", [&]() -> wholememory_error_code_t {
    using namespace wm;
    wholememory_tensor_t ts[4] = {csr_row_ptr_tensor, csr_col_ptr_tensor, output_csr_row_ptr_tensor, output_csr_col_ptr_tensor};
    for (auto t : ts) WM_REQUIRE_LIVE(t);
    /* argument checks in the reference's order and with its codes (graph_ops/csr_add_self_loop.cpp:26-87): rank and dtype
     * of each tensor in turn (INVALID_INPUT), then "views as an array" for each (LOGIC_ERROR) ... */
    for (auto t : ts) {
      auto* d = wholememory_tensor_get_tensor_description(t);
      if (d->dim != 1 || d->dtype != WHOLEMEMORY_DT_INT) {
        WM_ERROR("csr_add_self_loop: all four tensors should be 1D int tensors.");
        return WHOLEMEMORY_INVALID_INPUT;
      }
    }
    for (auto t : ts)
      if (!views_as_array(t)) {
        WM_ERROR("csr_add_self_loop: tensor convert to array failed.");
        return WHOLEMEMORY_LOGIC_ERROR;
      }
    /* ... then the point where the reference dispatches to its GPU translation unit */
    require_cuda("csr_add_self_loop");
    auto* rd = wholememory_tensor_get_tensor_description(csr_row_ptr_tensor);
    auto* cd = wholememory_tensor_get_tensor_description(csr_col_ptr_tensor);
    auto* od = wholememory_tensor_get_tensor_description(output_csr_row_ptr_tensor);
    auto* oc = wholememory_tensor_get_tensor_description(output_csr_col_ptr_tensor);
    const int64_t rows = rd->sizes[0] - 1;
    if (od->sizes[0] != rd->sizes[0] || oc->sizes[0] != cd->sizes[0] + rows) {
      WM_ERROR("csr_add_self_loop: output sizes must be rows+1 and edges+rows.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (rows <= 0) return WHOLEMEMORY_SUCCESS;
    add_self_loop_kernel<<<(unsigned)rows, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const int*>(wholememory_tensor_get_data_pointer(csr_row_ptr_tensor)),
      static_cast<const int*>(wholememory_tensor_get_data_pointer(csr_col_ptr_tensor)),
      static_cast<int*>(wholememory_tensor_get_data_pointer(output_csr_row_ptr_tensor)),
      static_cast<int*>(wholememory_tensor_get_data_pointer(output_csr_col_ptr_tensor)));
    WM_CUDA(cudaGetLastError());
    return WHOLEMEMORY_SUCCESS;
  });
}

} /* extern "C" */
