/*
 * Weighted CSR neighbor sampling without replacement (A-Res: keep the k largest keys log2(u)/w).
 * SURVEY 8(f) rank 4.  Replaces reference cpp/src/wholegraph_ops/weighted_sample_without_replacement_func.cuh
 * (key generation :45-63, fused block kernel :219-297, keys + segmented sort path for large k :501-600, driver :378-679)
 * and the host replay helper cpp/src/wholegraph_ops/raft_random_gen.cu:73-108.
 *
 * Same sample for the same (seed, inputs) as the reference, as far as the random stream is pinned (it is RAFT's
 * PCGenerator, restated: see sample_common.cuh): reference thread t of a B-thread CTA (B = 128 for k <= 256, else 256)
 * owns generator subsequence center*B + t and draws the keys of neighbours t, t+B, t+2B, ... in that order; every key
 * costs one float draw plus one-or-more 64-bit draws.
 *
 * Design: ONE WARP per center node (the reference uses a 128/256-thread CTA + RAFT warpsort::block_sort in shared memory).
 *  - k <= 32: lane l replays reference threads l, l+32, ...; the running top-k lives in registers, one (key, index) pair per
 *    lane, kept sorted; a chunk of 32 fresh keys is merged with a bitonic network ONLY when it contains a key above the
 *    current k-th best (one ballot decides), so most chunks of a high-degree node cost a compare and a vote.
 *  - k  > 32: keys of every over-full node go to a scratch array, cub::DeviceSegmentedSort orders each segment, the
 *    first k of each segment are emitted (same two-pass shape as the reference's large-k path).
 * Ties between equal keys are broken by the smaller neighbour index (the reference leaves them unspecified; its tests
 * compare per-node sorted outputs, cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:785).
 */
#include "sample_common.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <cfloat>
#include <cmath>

namespace wm {
namespace {

/* gen_key_from_weight, weighted_sample_without_replacement_func.cuh:45-63 */
template <typename WeightT>
__host__ __device__ __forceinline__ float key_from_weight(WeightT weight, pcg32& rng)
{
  float u = (float)(rng.next_u32() >> 8) / 16777216.0f; /* RAFT next(float&) */
  u       = -(0.5f + 0.5f * u);
  uint64_t r2 = 0;
  int rounds  = -1;
  do {
    uint64_t lo = rng.next_u32();
    uint64_t hi = rng.next_u32();
    r2          = lo | (hi << 32);
    ++rounds;
  } while (r2 == 0);
#ifdef __CUDA_ARCH__
  int one_bit = __clzll((long long)r2) + rounds * 64;
#else
  int one_bit = __builtin_clzll(r2) + rounds * 64;
#endif
  u *= exp2f(-(float)one_bit);
  return (log1pf(u) / logf(2.0f)) * (1.0f / (float)weight);
}

template <typename WeightT>
__device__ __forceinline__ WeightT load_weight(const table_ref& w, int64_t off_bytes, int64_t edge)
{
  return *reinterpret_cast<const WeightT*>(resolve_table_byte(w, (uint64_t)(off_bytes + edge * (int64_t)sizeof(WeightT))));
}

/* (key, idx) ordering: larger key first, then smaller index */
__device__ __forceinline__ bool better(float ka, int ia, float kb, int ib) { return ka > kb || (ka == kb && ia < ib); }

/* sort 32 (key, idx) pairs across the warp, best first */
__device__ __forceinline__ void warp_sort_desc(float& key, int& idx, int lane)
{
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      float ok      = __shfl_xor_sync(0xffffffffu, key, stride);
      int oi        = __shfl_xor_sync(0xffffffffu, idx, stride);
      bool up       = (lane & size) == 0;          /* this block sorts best-first, the next one worst-first */
      bool low_half = (lane & stride) == 0;        /* lower lane of the pair keeps the better element when `up` */
      bool other_better = better(ok, oi, key, idx);
      bool take         = (low_half == up) ? other_better : !other_better;
      if (take) {
        key = ok;
        idx = oi;
      }
    }
  }
}

/* merge step of a bitonic sequence (32 elements), best first */
__device__ __forceinline__ void warp_bitonic_merge_desc(float& key, int& idx, int lane)
{
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) {
    float ok          = __shfl_xor_sync(0xffffffffu, key, stride);
    int oi            = __shfl_xor_sync(0xffffffffu, idx, stride);
    bool low_half     = (lane & stride) == 0;
    bool other_better = better(ok, oi, key, idx);
    if (low_half ? other_better : !other_better) {
      key = ok;
      idx = oi;
    }
  }
}

constexpr int kWarps = 4;

template <typename IdT, typename ColT, typename WeightT>
__global__ void __launch_bounds__(kWarps * 32) weighted_topk_kernel(csr_ref g,
                                                                  table_ref weights,
                                                                  int64_t weight_off_bytes,
                                                                  const IdT* __restrict__ centers,
                                                                  int n,
                                                                  int k,
                                                                  int ref_block,
                                                                  uint64_t seed,
                                                                  const int* __restrict__ offsets,
                                                                  ColT* __restrict__ out_dst,
                                                                  int* __restrict__ out_center_lid,
                                                                  int64_t* __restrict__ out_edge_gid)
{
  const int lane = threadIdx.x & 31;
  const int c    = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (c >= n) return;
  const int64_t node = (int64_t)centers[c];
  int64_t start = 0, row_end = 0;
  node_bounds(g, c, n, node, &start, &row_end);
  const int N = (int)(row_end - start);
  if (N <= 0) return;
  const int off = offsets[c];
  if (k <= 0 || N <= k) {
    for (int s = lane; s < N; s += 32) {
      out_dst[off + s] = load_col<ColT>(g, start + s);
      if (out_center_lid) out_center_lid[off + s] = c;
      if (out_edge_gid) out_edge_gid[off + s] = start + s;
    }
    return;
  }
  /* running top-32 (only the first k matter), best first across lanes */
  float top_key = -FLT_MAX;
  int top_idx   = 0x7fffffff;
  for (int j = 0; j < ref_block / 32; ++j) {
    const int t = lane + 32 * j; /* reference thread replayed by this lane in this pass */
    pcg32 rng;
    rng.init(seed, (uint64_t)((int64_t)c * ref_block + t));
    for (int base = 32 * j; base < N; base += ref_block) { /* warp-uniform trip count */
      const int idx = base + lane;
      float key     = -FLT_MAX;
      int kidx      = 0x7fffffff;
      if (idx < N) {
        key  = key_from_weight<WeightT>(load_weight<WeightT>(weights, weight_off_bytes, start + idx), rng);
        kidx = idx;
      }
      /* k-th best so far sits in lane k-1 */
      const float thr_key = __shfl_sync(0xffffffffu, top_key, k - 1);
      const int thr_idx   = __shfl_sync(0xffffffffu, top_idx, k - 1);
      if (__ballot_sync(0xffffffffu, better(key, kidx, thr_key, thr_idx)) == 0) continue;
      warp_sort_desc(key, kidx, lane);
      /* best-32 of the union: elementwise best of (top best-first, chunk worst-first) is bitonic */
      const float rk = __shfl_sync(0xffffffffu, key, 31 - lane);
      const int ri   = __shfl_sync(0xffffffffu, kidx, 31 - lane);
      if (better(rk, ri, top_key, top_idx)) {
        top_key = rk;
        top_idx = ri;
      }
      warp_bitonic_merge_desc(top_key, top_idx, lane);
    }
  }
  if (lane < k) {
    out_dst[off + lane] = load_col<ColT>(g, start + top_idx);
    if (out_center_lid) out_center_lid[off + lane] = c;
    if (out_edge_gid) out_edge_gid[off + lane] = start + top_idx;
  }
}

/* ---- k > 32: keys to scratch, segmented sort, emit ---- */
template <typename IdT>
__global__ void overfull_degree_kernel(csr_ref g, const IdT* __restrict__ centers, int n, int k, int* __restrict__ seg_len)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int len = 0;
  if (i < n) {
    int64_t b = 0, e = 0;
    node_bounds(g, i, n, (int64_t)centers[i], &b, &e);
    int deg = (int)(e - b);
    len     = deg > k ? deg : 0; /* only over-full nodes need keys */
  }
  seg_len[i] = len;
}

template <typename IdT, typename WeightT>
__global__ void __launch_bounds__(kWarps * 32) weighted_keys_kernel(csr_ref g,
                                                                  table_ref weights,
                                                                  int64_t weight_off_bytes,
                                                                  const IdT* __restrict__ centers,
                                                                  int n,
                                                                  int k,
                                                                  int ref_block,
                                                                  uint64_t seed,
                                                                  const int* __restrict__ seg_off,
                                                                  float* __restrict__ keys,
                                                                  int* __restrict__ idxs)
{
  const int lane = threadIdx.x & 31;
  const int c    = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (c >= n) return;
  int64_t start = 0, row_end = 0;
  node_bounds(g, c, n, (int64_t)centers[c], &start, &row_end);
  const int N = (int)(row_end - start);
  if (N <= k) return;
  const int so = seg_off[c];
  for (int t = lane; t < ref_block; t += 32) {
    pcg32 rng;
    rng.init(seed, (uint64_t)((int64_t)c * ref_block + t));
    for (int idx = t; idx < N; idx += ref_block) {
      keys[so + idx] = key_from_weight<WeightT>(load_weight<WeightT>(weights, weight_off_bytes, start + idx), rng);
      idxs[so + idx] = idx;
    }
  }
}

template <typename IdT, typename ColT>
__global__ void __launch_bounds__(kWarps * 32) weighted_emit_kernel(csr_ref g,
                                                                  const IdT* __restrict__ centers,
                                                                  int n,
                                                                  int k,
                                                                  const int* __restrict__ offsets,
                                                                  const int* __restrict__ seg_off,
                                                                  const int* __restrict__ sorted_idx,
                                                                  ColT* __restrict__ out_dst,
                                                                  int* __restrict__ out_center_lid,
                                                                  int64_t* __restrict__ out_edge_gid)
{
  const int lane = threadIdx.x & 31;
  const int c    = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (c >= n) return;
  int64_t start = 0, row_end = 0;
  node_bounds(g, c, n, (int64_t)centers[c], &start, &row_end);
  const int N = (int)(row_end - start);
  if (N <= 0) return;
  const int off   = offsets[c];
  const bool all  = k <= 0 || N <= k;
  const int count = all ? N : k;
  for (int s = lane; s < count; s += 32) {
    const int nb     = all ? s : sorted_idx[seg_off[c] + s];
    out_dst[off + s] = load_col<ColT>(g, start + nb);
    if (out_center_lid) out_center_lid[off + s] = c;
    if (out_edge_gid) out_edge_gid[off + s] = start + nb;
  }
}

template <typename IdT, typename ColT, typename WeightT>
void run_weighted(const csr_ref& g, const table_ref& weights, int64_t weight_off_bytes, const void* centers, int n, int k, uint64_t seed,
                  int* offsets, wholememory_dtype_t col_dtype, void* dst_ctx, void* lid_ctx, void* gid_ctx, wholememory_env_func_t* env,
                  cudaStream_t s)
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
  read_back_sync(&total, offsets + n, sizeof(int), s);
  ColT* out_dst    = static_cast<ColT*>(output_alloc(env, dst_ctx, (size_t)total, col_dtype));
  int* out_lid     = lid_ctx ? static_cast<int*>(output_alloc(env, lid_ctx, (size_t)total, WHOLEMEMORY_DT_INT)) : nullptr;
  int64_t* out_gid = gid_ctx ? static_cast<int64_t*>(output_alloc(env, gid_ctx, (size_t)total, WHOLEMEMORY_DT_INT64)) : nullptr;
  if (n == 0 || total == 0) return;
  const int ref_block = k > 256 ? 256 : 128; /* reference func.cuh:532 / :626 (RAFT kMaxCapacity = 256) */
  const int grid      = (n + kWarps - 1) / kWarps;
  if (k <= 32) {
    weighted_topk_kernel<IdT, ColT, WeightT>
      <<<grid, kWarps * 32, 0, s>>>(g, weights, weight_off_bytes, cen, n, k, ref_block, seed, offsets, out_dst, out_lid, out_gid);
    WM_CUDA(cudaGetLastError());
    WM_CUDA(cudaStreamSynchronize(s));
    return;
  }
  /* large k: two passes */
  temp_buffer seg_len_b(env), seg_off_b(env), keys_a(env), keys_b(env), idx_a(env), idx_b(env), cub2_b(env);
  int* seg_len = static_cast<int*>(seg_len_b.device((size_t)n + 1, WHOLEMEMORY_DT_INT));
  int* seg_off = static_cast<int*>(seg_off_b.device((size_t)n + 1, WHOLEMEMORY_DT_INT));
  overfull_degree_kernel<IdT><<<(n + 1 + 127) / 128, 128, 0, s>>>(g, cen, n, k, seg_len);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, seg_len, seg_off, n + 1, s);
  int nkeys = 0;
  read_back_sync(&nkeys, seg_off + n, sizeof(int), s);
  int* sorted_idx = nullptr;
  if (nkeys > 0) {
    float* ka = static_cast<float*>(keys_a.device((size_t)nkeys, WHOLEMEMORY_DT_FLOAT));
    float* kb = static_cast<float*>(keys_b.device((size_t)nkeys, WHOLEMEMORY_DT_FLOAT));
    int* ia   = static_cast<int*>(idx_a.device((size_t)nkeys, WHOLEMEMORY_DT_INT));
    int* ib   = static_cast<int*>(idx_b.device((size_t)nkeys, WHOLEMEMORY_DT_INT));
    weighted_keys_kernel<IdT, WeightT><<<grid, kWarps * 32, 0, s>>>(g, weights, weight_off_bytes, cen, n, k, ref_block, seed, seg_off, ka, ia);
    size_t sort_bytes = 0;
    cub::DeviceSegmentedSort::StableSortPairsDescending(nullptr, sort_bytes, ka, kb, ia, ib, nkeys, n, seg_off, seg_off + 1, s);
    void* sort_tmp = cub2_b.device(sort_bytes, WHOLEMEMORY_DT_INT8);
    cub::DeviceSegmentedSort::StableSortPairsDescending(sort_tmp, sort_bytes, ka, kb, ia, ib, nkeys, n, seg_off, seg_off + 1, s);
    sorted_idx = ib;
  }
  weighted_emit_kernel<IdT, ColT><<<grid, kWarps * 32, 0, s>>>(g, cen, n, k, offsets, seg_off, sorted_idx, out_dst, out_lid, out_gid);
  WM_CUDA(cudaGetLastError());
  WM_CUDA(cudaStreamSynchronize(s));
}

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(wholememory_tensor_t wm_csr_row_ptr_tensor,
                                                                            wholememory_tensor_t wm_csr_col_ptr_tensor,
                                                                            wholememory_tensor_t wm_csr_weight_ptr_tensor,
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
  return wm::guarded("wholegraph_csr_weighted_sample_without_replacement", [&]() -> wholememory_error_code_t {
    using namespace wm;
    for (auto t : {wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor, center_nodes_tensor, output_sample_offset_tensor})
      WM_REQUIRE_LIVE(t);
    /* argument checks in the reference's order and with its codes (weighted_sample_without_replacement.cpp:74-126) ... */
    for (auto t : {wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor})
      if (!is_1d(t)) {
        WM_ERROR("wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor and wm_csr_weight_ptr_tensor should be 1D tensors.");
        return WHOLEMEMORY_INVALID_INPUT;
      }
    for (auto t : {wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor})
      if (!views_as_array(t)) {
        WM_ERROR("Input CSR tensor convert to array failed.");
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
    require_cuda("weighted neighbor sampling");
    auto rd = *wholememory_tensor_get_tensor_description(wm_csr_row_ptr_tensor);
    auto cd = *wholememory_tensor_get_tensor_description(wm_csr_col_ptr_tensor);
    auto wd = *wholememory_tensor_get_tensor_description(wm_csr_weight_ptr_tensor);
    auto nd = *wholememory_tensor_get_tensor_description(center_nodes_tensor);
    auto od = *wholememory_tensor_get_tensor_description(output_sample_offset_tensor);
    WM_EXPECT(rd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "wm_csr_row_ptr dtype must be int64");
    WM_EXPECT(od.dtype == WHOLEMEMORY_DT_INT, WHOLEMEMORY_LOGIC_ERROR, "output_sample_offset dtype must be int32");
    WM_EXPECT(cd.dtype == WHOLEMEMORY_DT_INT || cd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "col dtype must be int32/int64");
    WM_EXPECT(nd.dtype == WHOLEMEMORY_DT_INT || nd.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR, "center dtype must be int32/int64");
    WM_EXPECT(wd.dtype == WHOLEMEMORY_DT_FLOAT || wd.dtype == WHOLEMEMORY_DT_DOUBLE, WHOLEMEMORY_LOGIC_ERROR, "weight dtype must be float/double");
    WM_EXPECT(wd.sizes[0] == cd.sizes[0], WHOLEMEMORY_INVALID_INPUT, "one weight per edge expected");
    WM_EXPECT(od.sizes[0] == nd.sizes[0] + 1, WHOLEMEMORY_INVALID_INPUT, "output_sample_offset must have center_count + 1 entries");
    WM_EXPECT(nd.sizes[0] < ((int64_t)1 << 31) - 1, WHOLEMEMORY_INVALID_VALUE, "too many center nodes");
    for (auto t : {wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor})
      WM_EXPECT(!t->is_wm || handle_is_addressable(t->handle), WHOLEMEMORY_NOT_IMPLEMENTED,
                "weighted sampling needs peer-addressable CSR memory (no bucket-exchange variant; the reference has none either)");
    csr_ref g{};
    g.have_col             = 1;
    g.row_ptr              = make_table_ref(wm_csr_row_ptr_tensor);
    g.col                  = make_table_ref(wm_csr_col_ptr_tensor);
    g.row_ptr_offset_bytes = rd.storage_offset * 8;
    g.col_offset_bytes     = cd.storage_offset * (int64_t)wholememory_dtype_get_element_size(cd.dtype);
    table_ref w            = make_table_ref(wm_csr_weight_ptr_tensor);
    const int64_t w_off    = wd.storage_offset * (int64_t)wholememory_dtype_get_element_size(wd.dtype);
    const void* centers    = wholememory_tensor_get_data_pointer(center_nodes_tensor);
    int* offsets           = static_cast<int*>(wholememory_tensor_get_data_pointer(output_sample_offset_tensor));
    WM_EXPECT(offsets != nullptr && (centers != nullptr || nd.sizes[0] == 0), WHOLEMEMORY_INVALID_INPUT, "null center / offset pointer");
    auto s      = static_cast<cudaStream_t>(stream);
    const int n = (int)nd.sizes[0];
    const bool id64 = nd.dtype == WHOLEMEMORY_DT_INT64, col64 = cd.dtype == WHOLEMEMORY_DT_INT64, wf = wd.dtype == WHOLEMEMORY_DT_FLOAT;
#define WM_RUNW(IdT, ColT, WT)                                                                                          \
  run_weighted<IdT, ColT, WT>(g, w, w_off, centers, n, max_sample_count, random_seed, offsets, cd.dtype,                 \
                              output_dest_memory_context, output_center_localid_memory_context,                          \
                              output_edge_gid_memory_context, p_env_fns, s)
    if (id64 && col64) { if (wf) WM_RUNW(int64_t, int64_t, float); else WM_RUNW(int64_t, int64_t, double); }
    else if (id64) { if (wf) WM_RUNW(int64_t, int32_t, float); else WM_RUNW(int64_t, int32_t, double); }
    else if (col64) { if (wf) WM_RUNW(int32_t, int64_t, float); else WM_RUNW(int32_t, int64_t, double); }
    else { if (wf) WM_RUNW(int32_t, int32_t, float); else WM_RUNW(int32_t, int32_t, double); }
#undef WM_RUNW
    return WHOLEMEMORY_SUCCESS;
  });
}

/* host replay of the key stream for weight 1 (reference raft_random_gen.cu:73-108): output[i] = log2(u_i), u in (0,1) */
wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(int64_t random_seed, int64_t subsequence, wholememory_tensor_t output)
{
  WM_REQUIRE_LIVE(output);
  auto d = *wholememory_tensor_get_tensor_description(output);
  if (d.dim != 1) {
    WM_ERROR("output should be 1D tensor.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (d.dtype != WHOLEMEMORY_DT_FLOAT) {
    WM_ERROR("output should be float.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  float* p = static_cast<float*>(wholememory_tensor_get_data_pointer(output));
  wm::pcg32 rng;
  rng.init((uint64_t)random_seed, (uint64_t)subsequence);
  /* The reference's HOST replay is not the device formula: same draws, but the last step is double arithmetic,
   * log1p(u) / log(2.0) rounded once to float (raft_random_gen.cu:85-117), where the device kernels compute
   * log1pf(u) / logf(2.0) in float (func.cuh:45-63).  About a quarter of the values differ by one ulp, so the host
   * function follows the host code (pinned by tests/test_ref_host_random.py against the reference source itself). */
  for (int64_t i = 0; i < d.sizes[0]; ++i) {
    float u = (float)(rng.next_u32() >> 8) / 16777216.0f;
    u       = (float)-(0.5 + 0.5 * (double)u);
    uint64_t r2 = 0;
    int rounds  = -1;
    do {
      uint64_t lo = rng.next_u32();
      uint64_t hi = rng.next_u32();
      r2          = lo | (hi << 32);
      ++rounds;
    } while (r2 == 0);
    const int one_bit = __builtin_clzll(r2) + rounds * 64;
    u                 = (float)((double)u * pow(2.0, (double)-one_bit));
    p[i]              = (float)(log1p((double)u) / log(2.0));
  }
  return WHOLEMEMORY_SUCCESS;
}

} /* extern "C" */
