/*
 * Unweighted CSR neighbor sampling without replacement, reading row_ptr / col_idx out of
 * WholeMemory through the same owner-resolve + peer-load path as the gather kernel.
 *
 * Replaces reference cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh
 * (count :39-59, sampler :126-282, large-k :61-124, driver :284-475) and sample_comm.cuh:24-58.
 * Same outputs for the same (seed, inputs):
 *   offsets = exclusive scan of min(deg, k); deg <= k (or k <= 0) => all neighbours in CSR order;
 *   otherwise a[i] = Q[r[i]], Q[r[i]] = Q[N-i-1] with r[i] = draw(i) % (N - i)   (the recurrence the
 *   reference's radix-sort + pointer-jumping block kernel evaluates, and its CPU test restates:
 *   cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:306-321).
 *
 * Design: ONE WARP per center node instead of one CTA(32..256 threads) + cub::BlockRadixSort.  For
 * the fan-outs GNN loaders use (k <= 32) the whole recurrence lives in registers: lane j keeps the
 * j-th "back-fill" record (position, value) and a lookup Q[x] is one ballot + one shuffle; larger k
 * spill the records to a per-warp shared-memory list.  Four nodes per 128-thread CTA keep ~4x more
 * independent col_idx loads in flight per SM than the reference's block-per-node layout -- the op is
 * bound by NVLink/HBM transaction latency, not arithmetic.
 *
 * Random stream: PCG-XSH-RR 64/32 seeded the way RAFT's PCGenerator is (subsequence = center index *
 * BLOCK_DIM + t, BLOCK_DIM/ITEMS from the reference's (k-1)/32 table).  RAFT is an un-vendored
 * dependency (rapidsai/raft branch-24.12): restated from the published algorithm, parity UNPINNED.
 */
#include "sample_common.cuh"

#include <cub/device/device_scan.cuh>

#include <functional>

namespace wm {

/* exchange.cu */
void gather_by_exchange(wholememory_handle_t h,
                        const wholememory_matrix_description_t& table_desc,
                        const void* indices,
                        const wholememory_array_description_t& idx_desc,
                        void* output,
                        const wholememory_matrix_description_t& out_desc,
                        wholememory_env_func_t* env,
                        cudaStream_t stream,
                        int sms);

namespace {

/* (BLOCK_DIM, ITEMS_PER_THREAD) of the reference launch table (func.cuh:423-458) */
__host__ __device__ inline void reference_shape(int k, int* block_dim, int* items)
{
  const int f = (k - 1) / 32;
  int wc, it;
  if (f < 3) wc = 1;
  else if (f < 6) wc = 2;
  else if (f < 12) wc = 4;
  else wc = 8;
  if (f == 0) it = 1;
  else if (f == 1) it = 2;
  else if (f == 2) it = 3;
  else if (f == 3) it = 2;
  else if (f < 6) it = 3;
  else if (f < 8) it = 2;
  else if (f < 12) it = 3;
  else if (f < 16) it = 2;
  else if (f < 24) it = 3;
  else it = 4;
  *block_dim = wc * 32;
  *items     = it;
}

constexpr int kWarpsPerCta = 4;

/* exchange mode: ids[0..n) = centers, ids[n..2n) = centers + 1 (the two row_ptr entries every center needs) */
template <typename IdT>
__global__ void bounds_index_kernel(const IdT* __restrict__ centers, int n, int64_t* __restrict__ ids)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t c  = (int64_t)centers[i];
  ids[i]     = c;
  ids[n + i] = c + 1;
}

template <typename IdT, typename ColT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sample_kernel(csr_ref g,
                                                                  const IdT* __restrict__ centers,
                                                                  int n,
                                                                  int k,
                                                                  int ref_block_dim,
                                                                  uint64_t seed,
                                                                  const int* __restrict__ offsets,
                                                                  ColT* __restrict__ out_dst,
                                                                  int* __restrict__ out_center_lid,
                                                                  int64_t* __restrict__ out_edge_gid)
{
  extern __shared__ int smem[]; /* k > 32: per warp {pos[k], val[k], a[k]} */
  const int lane = threadIdx.x & 31;
  const int wid  = threadIdx.x >> 5;
  const int c    = blockIdx.x * kWarpsPerCta + wid;
  if (c >= n) return;
  const int64_t node = (int64_t)centers[c];
  int64_t start = 0, row_end = 0;
  node_bounds(g, c, n, node, &start, &row_end);
  const int N = (int)(row_end - start);
  if (N <= 0) return;
  const int off = offsets[c];

  if (k <= 0 || N <= k) { /* take every neighbour, CSR order */
    for (int s = lane; s < N; s += 32) {
      if (g.have_col) out_dst[off + s] = load_col<ColT>(g, start + s);
      if (out_center_lid) out_center_lid[off + s] = c;
      if (out_edge_gid) out_edge_gid[off + s] = start + s;
    }
    return;
  }

  const int M = k;
  if (M <= 32) {
    /* lane i draws r[i]: reference thread t = i (BLOCK_DIM = 32, ITEMS = 1), subsequence c*32 + t */
    int r = 0;
    if (lane < M) {
      pcg32 rng;
      rng.init(seed, (uint64_t)((int64_t)c * ref_block_dim + lane));
      r = rng.next_positive_int() % (N - lane);
    }
    int pos = -1, val = 0, a = 0; /* lane j: back-fill record of step j */
    for (int i = 0; i < M; ++i) {
      const int x1          = __shfl_sync(0xffffffffu, r, i);
      const int x2          = N - i - 1;
      const bool older      = lane < i;
      const unsigned hit1   = __ballot_sync(0xffffffffu, older && pos == x1);
      const unsigned hit2   = __ballot_sync(0xffffffffu, older && pos == x2);
      const int v1          = __shfl_sync(0xffffffffu, val, hit1 ? 31 - __clz(hit1) : 0);
      const int v2          = __shfl_sync(0xffffffffu, val, hit2 ? 31 - __clz(hit2) : 0);
      const int q1          = hit1 ? v1 : x1; /* Q[r[i]]   */
      const int q2          = hit2 ? v2 : x2; /* Q[N-i-1]  */
      if (lane == i) {
        a   = q1;
        pos = x1;
        val = q2;
      }
    }
    if (lane < M) {
      if (g.have_col) out_dst[off + lane] = load_col<ColT>(g, start + a);
      if (out_center_lid) out_center_lid[off + lane] = c;
      if (out_edge_gid) out_edge_gid[off + lane] = start + a;
    }
    return;
  }

  /* general k (33..1024): records in shared memory, 32-lane parallel scan per lookup */
  int* pos_s = smem + wid * 3 * M;
  int* val_s = pos_s + M;
  int* a_s   = val_s + M;
  for (int i = lane; i < M; i += 32) {
    /* position i = j*BLOCK_DIM + t is draw j of reference thread t */
    const int t = i % ref_block_dim, j = i / ref_block_dim;
    pcg32 rng;
    rng.init(seed, (uint64_t)((int64_t)c * ref_block_dim + t));
    int draw = 0;
    for (int d = 0; d <= j; ++d) draw = rng.next_positive_int();
    a_s[i] = draw % (N - i); /* r[i], replaced by a[i] below */
  }
  __syncwarp();
  for (int i = 0; i < M; ++i) {
    const int x1 = a_s[i];
    const int x2 = N - i - 1;
    int j1 = -1, j2 = -1;
    for (int j = lane; j < i; j += 32) {
      const int p = pos_s[j];
      if (p == x1) j1 = j;
      if (p == x2) j2 = j;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      j1 = max(j1, __shfl_xor_sync(0xffffffffu, j1, d));
      j2 = max(j2, __shfl_xor_sync(0xffffffffu, j2, d));
    }
    if (lane == 0) {
      const int q1 = j1 >= 0 ? val_s[j1] : x1;
      const int q2 = j2 >= 0 ? val_s[j2] : x2;
      a_s[i]       = q1;
      pos_s[i]     = x1;
      val_s[i]     = q2;
    }
    __syncwarp();
  }
  for (int i = lane; i < M; i += 32) {
    const int a      = a_s[i];
    if (g.have_col) out_dst[off + i] = load_col<ColT>(g, start + a);
    if (out_center_lid) out_center_lid[off + i] = c;
    if (out_edge_gid) out_edge_gid[off + i] = start + a;
  }
}

/* k > 1024: reservoir-style selection of the reference's large_sample_kernel (func.cuh:61-124);
 * result depends on atomicMax outcomes only, so it is deterministic. One CTA of 32 threads per node. */
template <typename IdT, typename ColT>
__global__ void large_sample_kernel(csr_ref g,
                                    const IdT* __restrict__ centers,
                                    int n,
                                    int k,
                                    uint64_t seed,
                                    const int* __restrict__ offsets,
                                    ColT* __restrict__ out_dst,
                                    int* __restrict__ out_center_lid,
                                    int64_t* __restrict__ out_edge_gid,
                                    int* __restrict__ scratch /* total samples ints */)
{
  const int c = blockIdx.x;
  if (c >= n) return;
  pcg32 rng;
  rng.init(seed, (uint64_t)(threadIdx.x + (int64_t)blockIdx.x * blockDim.x));
  const int64_t node = (int64_t)centers[c];
  int64_t start = 0, row_end = 0;
  node_bounds(g, c, n, node, &start, &row_end);
  const int N   = (int)(row_end - start);
  const int off = offsets[c];
  if (N <= k) {
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
      if (g.have_col) out_dst[off + s] = load_col<ColT>(g, start + s);
      if (out_center_lid) out_center_lid[off + s] = c;
      if (out_edge_gid) out_edge_gid[off + s] = start + s;
    }
    return;
  }
  int* slot = scratch + off;
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    slot[s] = s;
    if (out_center_lid) out_center_lid[off + s] = c;
  }
  __syncthreads();
  for (int idx = k + threadIdx.x; idx < N; idx += blockDim.x) {
    int rnd = rng.next_positive_int() % (idx + 1);
    if (rnd < k) atomicMax(slot + rnd, idx);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    int nb           = slot[s];
    if (g.have_col) out_dst[off + s] = load_col<ColT>(g, start + nb);
    if (out_edge_gid) out_edge_gid[off + s] = start + nb;
  }
}

template <typename IdT, typename ColT>
void run_sampler(const csr_ref& g, const void* centers, int n, int k, uint64_t seed, int* offsets, wholememory_dtype_t col_dtype,
                 void* dst_ctx, void* lid_ctx, void* gid_ctx, wholememory_env_func_t* env, cudaStream_t s,
                 const std::function<void(const int64_t* edge_ids, void* dst, int total)>& fetch_col = nullptr)
{
  const IdT* cen = static_cast<const IdT*>(centers);
  temp_buffer counts_b(env), cub_b(env);
  int* counts = static_cast<int*>(counts_b.device((size_t)n + 1, WHOLEMEMORY_DT_INT));
  sample_count_kernel<IdT><<<(n + 1 + 127) / 128, 128, 0, s>>>(g, cen, n, k, counts);
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, counts, offsets, n + 1, s);
  void* cub_tmp = cub_b.device(cub_bytes, WHOLEMEMORY_DT_INT8);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, offsets, n + 1, s);
  int total = 0;
  read_back_sync(&total, offsets + n, sizeof(int), s); /* the output allocation callback needs the size */

  ColT* out_dst   = static_cast<ColT*>(output_alloc(env, dst_ctx, (size_t)total, col_dtype));
  int* out_lid    = lid_ctx ? static_cast<int*>(output_alloc(env, lid_ctx, (size_t)total, WHOLEMEMORY_DT_INT)) : nullptr;
  int64_t* out_gid = gid_ctx ? static_cast<int64_t*>(output_alloc(env, gid_ctx, (size_t)total, WHOLEMEMORY_DT_INT64)) : nullptr;
  temp_buffer gid_tmp(env);
  if (!g.have_col && out_gid == nullptr) /* exchange mode always needs the edge ids: they drive the col_idx fetch */
    out_gid = static_cast<int64_t*>(gid_tmp.device((size_t)std::max(total, 1), WHOLEMEMORY_DT_INT64));
  if (n == 0 || total == 0) {
    if (fetch_col) fetch_col(out_gid, out_dst, 0); /* the exchange is collective: take part even with nothing to fetch */
    return;
  }

  if (k > 1024) {
    temp_buffer scratch_b(env);
    int* scratch = static_cast<int*>(scratch_b.device((size_t)total, WHOLEMEMORY_DT_INT));
    large_sample_kernel<IdT, ColT><<<n, 32, 0, s>>>(g, cen, n, k, seed, offsets, out_dst, out_lid, out_gid, scratch);
    WM_CUDA(cudaGetLastError());
    if (fetch_col) fetch_col(out_gid, out_dst, total);
    WM_CUDA(cudaStreamSynchronize(s));
    return;
  }
  int ref_block_dim = 32, ref_items = 1;
  if (k > 0) reference_shape(k, &ref_block_dim, &ref_items);
  size_t smem = k > 32 ? (size_t)kWarpsPerCta * 3 * k * sizeof(int) : 0;
  if (smem > 48 * 1024)
    WM_CUDA(cudaFuncSetAttribute(sample_kernel<IdT, ColT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_kernel<IdT, ColT><<<(n + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, smem, s>>>(
    g, cen, n, k, ref_block_dim, seed, offsets, out_dst, out_lid, out_gid);
  WM_CUDA(cudaGetLastError());
  if (fetch_col) fetch_col(out_gid, out_dst, total); /* exchange mode: col_idx[edge id] via the bucket exchange */
  WM_CUDA(cudaStreamSynchronize(s)); /* reference func.cuh:474; temporaries are released on return */
}

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement(wholememory_tensor_t wm_csr_row_ptr_tensor,
                                                                              wholememory_tensor_t wm_csr_col_ptr_tensor,
                                                                              wholememory_tensor_t center_nodes_tensor,
                                                                              int max_sample_count,
                                                                              wholememory_tensor_t output_sample_offset_tensor,
                                                                              void* output_dest_memory_context,
                                                                              void* output_center_localid_memory_context,
                                                                              void* output_edge_gid_memory_context,
                                                                              unsigned long long random_seed,
                                                                              wholememory_env_func_t* p_env_fns,
                                                                              void* stream)
{
  return wm::guarded("wholegraph_csr_unweighted_sample_without_replacement", [&]() -> wholememory_error_code_t {
    using namespace wm;
    for (auto t : {wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, center_nodes_tensor, output_sample_offset_tensor}) WM_REQUIRE_LIVE(t);
    /* argument checks in the reference's order and with its codes (unweighted_sample_without_replacement.cpp:64-111) ... */
    if (!is_1d(wm_csr_row_ptr_tensor)) {
      WM_ERROR("wm_csr_row_ptr_tensor should be 1D tensor.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (!is_1d(wm_csr_col_ptr_tensor)) {
      WM_ERROR("wm_csr_col_ptr_tensor should be 1D tensor.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (!views_as_array(wm_csr_row_ptr_tensor) || !views_as_array(wm_csr_col_ptr_tensor)) {
      WM_ERROR("Input wm_csr_row_ptr_tensor / wm_csr_col_ptr_tensor convert to array failed.");
      return WHOLEMEMORY_LOGIC_ERROR;
    }
    if (!is_1d(center_nodes_tensor)) {
      WM_ERROR("Input center_nodes_tensor should be 1D tensor");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (!views_as_array(center_nodes_tensor)) {
      WM_ERROR("Input center_nodes_tensor convert to array failed.");
      return WHOLEMEMORY_LOGIC_ERROR;
    }
    if (!is_1d(output_sample_offset_tensor)) {
      WM_ERROR("Output output_sample_offset_tensor should be 1D tensor.");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (!views_as_array(output_sample_offset_tensor)) {
      WM_ERROR("Output output_sample_offset_tensor convert to array failed.");
      return WHOLEMEMORY_LOGIC_ERROR;
    }
    /* ... then the point where the reference dispatches to its GPU translation unit */
    require_cuda("neighbor sampling");
    auto rd = *wholememory_tensor_get_tensor_description(wm_csr_row_ptr_tensor);
    auto cd = *wholememory_tensor_get_tensor_description(wm_csr_col_ptr_tensor);
    auto nd = *wholememory_tensor_get_tensor_description(center_nodes_tensor);
    auto od = *wholememory_tensor_get_tensor_description(output_sample_offset_tensor);
    /* dtype rules: reference unweighted_sample_without_replacement_func.cuh:304-315 */
    WM_EXPECT(rd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "wm_csr_row_ptr dtype must be int64, got %d", (int)rd.dtype);
    WM_EXPECT(od.dtype == WHOLEMEMORY_DT_INT, WHOLEMEMORY_LOGIC_ERROR, "output_sample_offset dtype must be int32, got %d", (int)od.dtype);
    WM_EXPECT(cd.dtype == WHOLEMEMORY_DT_INT || cd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "col dtype must be int32/int64");
    WM_EXPECT(nd.dtype == WHOLEMEMORY_DT_INT || nd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "center dtype must be int32/int64");
    WM_EXPECT(od.sizes[0] == nd.sizes[0] + 1, WHOLEMEMORY_INVALID_INPUT, "output_sample_offset must have center_count + 1 entries");
    WM_EXPECT(nd.sizes[0] < ((int64_t)1 << 31) - 1, WHOLEMEMORY_INVALID_VALUE, "too many center nodes");
    const void* centers = wholememory_tensor_get_data_pointer(center_nodes_tensor);
    int* offsets        = static_cast<int*>(wholememory_tensor_get_data_pointer(output_sample_offset_tensor));
    WM_EXPECT(offsets != nullptr && (centers != nullptr || nd.sizes[0] == 0), WHOLEMEMORY_INVALID_INPUT, "null center / offset pointer");
    auto s          = static_cast<cudaStream_t>(stream);
    const int n     = (int)nd.sizes[0];
    const bool id64 = nd.dtype == WHOLEMEMORY_DT_INT64, col64 = cd.dtype == WHOLEMEMORY_DT_INT64;
    const bool row_direct = !wm_csr_row_ptr_tensor->is_wm || handle_is_addressable(wm_csr_row_ptr_tensor->handle);
    const bool col_direct = !wm_csr_col_ptr_tensor->is_wm || handle_is_addressable(wm_csr_col_ptr_tensor->handle);
    csr_ref g{};
    g.have_col = col_direct ? 1 : 0;
    if (row_direct) {
      g.row_ptr              = make_table_ref(wm_csr_row_ptr_tensor);
      g.row_ptr_offset_bytes = rd.storage_offset * 8;
    }
    if (col_direct) {
      g.col              = make_table_ref(wm_csr_col_ptr_tensor);
      g.col_offset_bytes = cd.storage_offset * (int64_t)wholememory_dtype_get_element_size(cd.dtype);
    }
    /* DISTRIBUTED memory without peer mapping (reference ..._nccl_func.cuh:193-387): fetch the row_ptr pairs through the
     * bucket exchange first, sample edge ids locally, then fetch col_idx[edge id] through the exchange.  Collective. */
    temp_buffer bounds_ids(p_env_fns), bounds_vals(p_env_fns);
    if (!row_direct) {
      auto* ids  = static_cast<int64_t*>(bounds_ids.device((size_t)std::max(2 * n, 1), WHOLEMEMORY_DT_INT64));
      auto* vals = static_cast<int64_t*>(bounds_vals.device((size_t)std::max(2 * n, 1), WHOLEMEMORY_DT_INT64));
      if (n > 0) {
        if (id64) bounds_index_kernel<int64_t><<<(n + 255) / 256, 256, 0, s>>>(static_cast<const int64_t*>(centers), n, ids);
        else bounds_index_kernel<int32_t><<<(n + 255) / 256, 256, 0, s>>>(static_cast<const int32_t*>(centers), n, ids);
      }
      int64_t tsz[2] = {rd.sizes[0], 1}, osz[2] = {2 * (int64_t)n, 1};
      gather_by_exchange(wm_csr_row_ptr_tensor->handle, wholememory_create_matrix_desc(tsz, 1, rd.storage_offset, WHOLEMEMORY_DT_INT64), ids,
                         wholememory_create_array_desc(2 * (int64_t)n, 0, WHOLEMEMORY_DT_INT64), vals,
                         wholememory_create_matrix_desc(osz, 1, 0, WHOLEMEMORY_DT_INT64), p_env_fns, s, -1);
      g.pre_bounds = vals;
    }
    std::function<void(const int64_t*, void*, int)> fetch_col;
    if (!col_direct) {
      fetch_col = [&](const int64_t* edge_ids, void* dst, int total) {
        int64_t tsz[2] = {cd.sizes[0], 1}, osz[2] = {(int64_t)total, 1};
        gather_by_exchange(wm_csr_col_ptr_tensor->handle, wholememory_create_matrix_desc(tsz, 1, cd.storage_offset, cd.dtype), edge_ids,
                           wholememory_create_array_desc((int64_t)total, 0, WHOLEMEMORY_DT_INT64), dst,
                           wholememory_create_matrix_desc(osz, 1, 0, cd.dtype), p_env_fns, s, -1);
      };
    }
#define WM_RUN(IdT, ColT)                                                                                                   \
  run_sampler<IdT, ColT>(g, centers, n, max_sample_count, random_seed, offsets, cd.dtype, output_dest_memory_context,       \
                         output_center_localid_memory_context, output_edge_gid_memory_context, p_env_fns, s, fetch_col)
    if (id64 && col64) WM_RUN(int64_t, int64_t);
    else if (id64) WM_RUN(int64_t, int32_t);
    else if (col64) WM_RUN(int32_t, int64_t);
    else WM_RUN(int32_t, int32_t);
#undef WM_RUN
    return WHOLEMEMORY_SUCCESS;
  });
}

/* host replay of the sampler's stream (reference raft_random_gen.cu:27-71) */
wholememory_error_code_t generate_random_positive_int_cpu(int64_t random_seed, int64_t subsequence, wholememory_tensor_t output)
{
  WM_REQUIRE_LIVE(output);
  auto d = *wholememory_tensor_get_tensor_description(output);
  if (d.dim != 1) {
    WM_ERROR("output should be 1D tensor.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (d.dtype != WHOLEMEMORY_DT_INT64 && d.dtype != WHOLEMEMORY_DT_INT) {
    WM_ERROR("output should be int64 or int32 tensor.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  void* p = wholememory_tensor_get_data_pointer(output);
  wm::pcg32 rng;
  rng.init((uint64_t)random_seed, (uint64_t)subsequence);
  for (int64_t i = 0; i < d.sizes[0]; ++i) {
    if (d.dtype == WHOLEMEMORY_DT_INT) static_cast<int32_t*>(p)[i] = rng.next_positive_int();
    else static_cast<int64_t*>(p)[i] = rng.next_positive_int64();
  }
  return WHOLEMEMORY_SUCCESS;
}

} /* extern "C" */
