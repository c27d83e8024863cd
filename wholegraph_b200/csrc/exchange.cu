/*
 * Bucket exchange for DISTRIBUTED WholeMemory whose shards are NOT peer-addressable, and for the
 * gradient path of trainable embeddings.
 *
 * Replaces reference cpp/src/wholememory_ops/functions/bucket_ids_func.cu:51-139,
 * exchange_ids_nccl_func.cu:42-226, exchange_embeddings_nccl_func.cu:32-74 and the drivers
 * gather_op_impl_nccl.cu:34-182 / scatter_op_impl_nccl.cu:34-181.
 *
 * Differences from the reference pipeline:
 *  - indices are grouped by owner with a hand-written STABLE 3-kernel counting partition
 *    (<= 17 buckets: ranks + "discard") instead of a full 32/64-bit cub radix sort + iota; the
 *    permutation is deterministic, so duplicate-gradient sums downstream are reproducible.
 *  - one host synchronisation (bucket totals) instead of two; the count all-to-all travels over
 *    the AF_UNIX bootstrap instead of pinned bounce buffers + NCCL.
 *  - negative (and out-of-range) indices fall into the discard bucket and are never sent; their
 *    output rows are left untouched, matching the mapped path's contract.
 */
#include "exchange.hpp"

#include <algorithm>

namespace wm {

namespace {

constexpr int kWords       = 5; /* 5 x 4 sixteen-bit counters = 20 buckets >= 16 ranks + discard */
constexpr int kTileThreads = 256;
constexpr int kItems       = 4;
constexpr int kTile        = kTileThreads * kItems;

struct owner_map {
  int nranks;
  bool regular;
  int64_t rows_per_rank;
  int64_t first_row[kMaxInlineRanks + 1];
};

__device__ __forceinline__ int owner_of(const owner_map& m, int64_t idx)
{
  if (idx < 0 || idx >= m.first_row[m.nranks]) return m.nranks; /* discard */
  if (m.regular) return (int)(idx / m.rows_per_rank);
  int o = 0;
#pragma unroll 1
  for (int r = 1; r < m.nranks; ++r)
    if (idx >= m.first_row[r]) o = r;
  return o;
}

struct packed {
  uint64_t w[kWords];
};
__device__ __forceinline__ void packed_inc(packed& p, int b)
{
  uint64_t inc = 1ull << ((b & 3) * 16);
#pragma unroll
  for (int j = 0; j < kWords; ++j) p.w[j] += (j == (b >> 2)) ? inc : 0ull;
}
__device__ __forceinline__ uint32_t packed_get(const packed& p, int b)
{
  uint64_t w = 0;
#pragma unroll
  for (int j = 0; j < kWords; ++j) w = (j == (b >> 2)) ? p.w[j] : w;
  return (uint32_t)(w >> ((b & 3) * 16)) & 0xffffu;
}

/* A: per-tile histogram (order independent -> shared atomics are fine) */
template <typename IdxT>
__global__ void __launch_bounds__(kTileThreads) tile_hist_kernel(const IdxT* __restrict__ idx, int64_t n, owner_map m, uint32_t* __restrict__ tile_hist)
{
  __shared__ uint32_t hist[kMaxInlineRanks + 1];
  if (threadIdx.x <= m.nranks) hist[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int64_t i = base + k;
    if (i < n) atomicAdd(&hist[owner_of(m, (int64_t)idx[i])], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= m.nranks) tile_hist[(int64_t)blockIdx.x * (m.nranks + 1) + threadIdx.x] = hist[threadIdx.x];
}

/* B: column-wise exclusive scan over tiles; bucket totals and bucket bases */
__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ tile_hist,
                                                        uint32_t* __restrict__ tile_base,
                                                        int64_t ntiles,
                                                        int nb,
                                                        int64_t* __restrict__ totals /* [nb] */,
                                                        int64_t* __restrict__ bucket_base /* [nb] */)
{
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int b = 0; b < nb; ++b) {
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < ntiles; t0 += blockDim.x) {
      int64_t t  = t0 + threadIdx.x;
      uint32_t v = t < ntiles ? tile_hist[t * nb + b] : 0u;
      uint32_t s = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += o;
      }
      if (lane == 31) warp_sum[wid] = s;
      __syncthreads();
      if (wid == 0) {
        uint32_t ws_ = warp_sum[lane], acc = ws_;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, acc, d);
          if (lane >= d) acc += o;
        }
        warp_sum[lane] = acc - ws_; /* exclusive */
      }
      __syncthreads();
      uint32_t carry = carry_s;
      uint32_t excl  = carry + warp_sum[wid] + (s - v);
      if (t < ntiles) tile_base[t * nb + b] = excl;
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
      __syncthreads();
    }
    if (threadIdx.x == 0) totals[b] = (int64_t)carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int b = 0; b < nb; ++b) {
      bucket_base[b] = acc;
      acc += totals[b];
    }
  }
}

/* C: stable placement */
template <typename IdxT>
__global__ void __launch_bounds__(kTileThreads) tile_place_kernel(const IdxT* __restrict__ idx,
                                                                 int64_t n,
                                                                 owner_map m,
                                                                 const uint32_t* __restrict__ tile_base,
                                                                 const int64_t* __restrict__ bucket_base,
                                                                 IdxT* __restrict__ grouped_idx,
                                                                 int64_t* __restrict__ origin)
{
  __shared__ packed warp_tot[kTileThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nb   = m.nranks + 1;
  int64_t base   = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  IdxT val[kItems];
  int own[kItems];
  packed mine{};
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int64_t i = base + k;
    own[k]    = -1;
    if (i < n) {
      val[k] = idx[i];
      own[k] = owner_of(m, (int64_t)val[k]);
      packed_inc(mine, own[k]);
    }
  }
  /* block-wide exclusive scan of the packed histograms (thread order == index order) */
  packed incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int j = 0; j < kWords; ++j) {
      uint64_t o = __shfl_up_sync(0xffffffffu, incl.w[j], d);
      if (lane >= d) incl.w[j] += o;
    }
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  packed prefix{};
  for (int w = 0; w < wid; ++w) {
#pragma unroll
    for (int j = 0; j < kWords; ++j) prefix.w[j] += warp_tot[w].w[j];
  }
#pragma unroll
  for (int j = 0; j < kWords; ++j) prefix.w[j] += incl.w[j] - mine.w[j];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int b = own[k];
    if (b < 0) continue;
    uint32_t within = packed_get(prefix, b);
    packed_inc(prefix, b);
    if (b < m.nranks) {
      int64_t pos      = bucket_base[b] + tile_base[(int64_t)blockIdx.x * nb + b] + within;
      grouped_idx[pos] = val[k];
      origin[pos]      = base + k;
    }
  }
}

template <typename IdxT>
void run_partition(const void* idx, int64_t n, const owner_map& m, uint32_t* tile_hist, uint32_t* tile_base,
                   int64_t ntiles, int64_t* totals, int64_t* bucket_base, void* grouped, int64_t* origin, cudaStream_t s,
                   int64_t* host_totals, cudaEvent_t totals_ready)
{
  const int nb = m.nranks + 1;
  tile_hist_kernel<IdxT><<<(unsigned)ntiles, kTileThreads, 0, s>>>(static_cast<const IdxT*>(idx), n, m, tile_hist);
  tile_scan_kernel<<<1, 1024, 0, s>>>(tile_hist, tile_base, ntiles, nb, totals, bucket_base);
  WM_CUDA(cudaMemcpyAsync(host_totals, totals, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, s));
  WM_CUDA(cudaEventRecord(totals_ready, s));
  tile_place_kernel<IdxT><<<(unsigned)ntiles, kTileThreads, 0, s>>>(static_cast<const IdxT*>(idx), n, m, tile_base, bucket_base,
                                                                    static_cast<IdxT*>(grouped), origin);
  WM_CUDA(cudaGetLastError());
}

}  // namespace

exchange_plan::exchange_plan(wholememory_env_func_t* env)
  : grouped_idx(env), origin(env), recv_idx(env), scratch_hist(env), scratch_base(env), scratch_totals(env), host_totals(env)
{
}

void partition_by_owner(exchange_plan* p,
                        wholememory_comm_t comm,
                        const void* indices,
                        wholememory_dtype_t idx_dtype,
                        int64_t n,
                        const std::vector<int64_t>& first_row,
                        cudaStream_t stream)
{
  const int ws = comm->world_size;
  WM_EXPECT(ws <= kMaxInlineRanks, WHOLEMEMORY_NOT_SUPPORTED, "bucket exchange supports at most %d ranks", kMaxInlineRanks);
  WM_EXPECT(n < (int64_t)1 << 31, WHOLEMEMORY_INVALID_VALUE, "too many indices in one call (%ld)", (long)n);
  const bool idx64 = idx_dtype == WHOLEMEMORY_DT_INT64;
  WM_EXPECT(idx64 || idx_dtype == WHOLEMEMORY_DT_INT, WHOLEMEMORY_LOGIC_ERROR, "indices must be int32 or int64");
  p->idx_dtype = idx_dtype;
  p->n         = n;
  p->send_counts.assign(ws, 0);
  p->recv_counts.assign(ws, 0);

  owner_map m{};
  m.nranks        = ws;
  m.rows_per_rank = first_row[1] - first_row[0];
  for (int r = 0; r <= ws; ++r) m.first_row[r] = first_row[r];
  /* regular <=> boundaries are exactly min(r * rows_per_rank, total) */
  m.regular = m.rows_per_rank > 0;
  for (int r = 0; r <= ws && m.regular; ++r)
    if (first_row[r] != std::min<int64_t>((int64_t)r * m.rows_per_rank, first_row[ws])) m.regular = false;
  if (!m.regular) m.rows_per_rank = 1;

  const int nb = ws + 1;
  if (n > 0) {
    int64_t ntiles = (n + kTile - 1) / kTile;
    auto* hist     = static_cast<uint32_t*>(p->scratch_hist.device((size_t)ntiles * nb, WHOLEMEMORY_DT_INT));
    auto* tbase    = static_cast<uint32_t*>(p->scratch_base.device((size_t)ntiles * nb, WHOLEMEMORY_DT_INT));
    auto* totals   = static_cast<int64_t*>(p->scratch_totals.device((size_t)nb * 2, WHOLEMEMORY_DT_INT64));
    auto* htot     = static_cast<int64_t*>(p->host_totals.pinned((size_t)nb, WHOLEMEMORY_DT_INT64));
    p->grouped_idx.device((size_t)n, idx_dtype);
    p->origin.device((size_t)n, WHOLEMEMORY_DT_INT64);
    cudaEvent_t ev;
    WM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    try {
      if (idx64)
        run_partition<int64_t>(indices, n, m, hist, tbase, ntiles, totals, totals + nb, p->grouped_idx.ptr(),
                               static_cast<int64_t*>(p->origin.ptr()), stream, htot, ev);
      else
        run_partition<int32_t>(indices, n, m, hist, tbase, ntiles, totals, totals + nb, p->grouped_idx.ptr(),
                               static_cast<int64_t*>(p->origin.ptr()), stream, htot, ev);
      WM_CUDA(cudaEventSynchronize(ev)); /* the one host sync: totals must be known to size the exchange */
    } catch (...) {
      cudaEventDestroy(ev);
      throw;
    }
    cudaEventDestroy(ev);
    for (int r = 0; r < ws; ++r) p->send_counts[r] = htot[r];
  }
  p->n_send = 0;
  for (int r = 0; r < ws; ++r) p->n_send += p->send_counts[r];
}

void plan_exchange(exchange_plan* p,
                   wholememory_comm_t comm,
                   const void* indices,
                   wholememory_dtype_t idx_dtype,
                   int64_t n,
                   const std::vector<int64_t>& first_row,
                   cudaStream_t stream)
{
  partition_by_owner(p, comm, indices, idx_dtype, n, first_row, stream);
  const int ws     = comm->world_size;
  const bool idx64 = idx_dtype == WHOLEMEMORY_DT_INT64;
  {
    std::lock_guard<std::mutex> lk(comm->mu);
    comm->boot->alltoall(p->send_counts.data(), p->recv_counts.data(), sizeof(int64_t));
  }
  p->n_send = 0;
  p->n_recv = 0;
  for (int r = 0; r < ws; ++r) {
    p->n_send += p->send_counts[r];
    p->n_recv += p->recv_counts[r];
  }
  /* ship the grouped indices to their owners */
  const size_t isz = idx64 ? 8 : 4;
  p->recv_idx.device((size_t)std::max<int64_t>(p->n_recv, 1), idx_dtype);
  exchange_rows(*p, comm, p->grouped_idx.ptr(), p->recv_idx.ptr(), isz, /*to_owner=*/true, stream);
}

void exchange_rows(const exchange_plan& p,
                   wholememory_comm_t comm,
                   const void* send,
                   void* recv,
                   size_t row_bytes,
                   bool to_owner,
                   cudaStream_t stream)
{
  const int ws = comm->world_size;
  const auto& sc = to_owner ? p.send_counts : p.recv_counts;
  const auto& rc = to_owner ? p.recv_counts : p.send_counts;
  std::vector<size_t> sb(ws), sd(ws), rb(ws), rd(ws);
  size_t so = 0, ro = 0;
  for (int r = 0; r < ws; ++r) {
    sb[r] = (size_t)sc[r] * row_bytes;
    rb[r] = (size_t)rc[r] * row_bytes;
    sd[r] = so;
    rd[r] = ro;
    so += sb[r];
    ro += rb[r];
  }
  nccl_alltoallv_bytes(comm, send, sb.data(), sd.data(), recv, rb.data(), rd.data(), stream);
}

std::vector<int64_t> handle_first_rows(wholememory_handle_t h, size_t row_stride_bytes)
{
  const int ws = h->comm->world_size;
  std::vector<int64_t> fr(ws + 1);
  for (int r = 0; r <= ws; ++r) {
    WM_EXPECT(h->part_offsets[r] % row_stride_bytes == 0, WHOLEMEMORY_LOGIC_ERROR,
              "partition offset of rank %d (%zu) is not a multiple of the row stride (%zu bytes)", r, h->part_offsets[r], row_stride_bytes);
    fr[r] = (int64_t)(h->part_offsets[r] / row_stride_bytes);
  }
  return fr;
}

/* table_ref that lets GLOBAL row ids address only this rank's shard */
static table_ref local_shard_ref(wholememory_handle_t h)
{
  char* fake = static_cast<char*>(h->local_ptr) - h->part_offsets[h->comm->world_rank];
  return make_flat_table_ref(fake);
}

void gather_by_exchange(wholememory_handle_t h,
                        const wholememory_matrix_description_t& td,
                        const void* indices,
                        const wholememory_array_description_t& idx_desc,
                        void* output,
                        const wholememory_matrix_description_t& od,
                        wholememory_env_func_t* env,
                        cudaStream_t stream,
                        int sms)
{
  require_cuda("wholememory_gather (DISTRIBUTED)");
  /* only column windows of whole rows can be exchanged (reference gather_op_impl_nccl.cu:45-48) */
  WM_EXPECT(td.storage_offset >= 0 && td.storage_offset + td.sizes[1] <= td.stride, WHOLEMEMORY_INVALID_INPUT,
            "DISTRIBUTED gather needs a tensor that starts at row 0");
  WM_EXPECT(od.sizes[0] == idx_desc.size, WHOLEMEMORY_LOGIC_ERROR, "output rows=%ld but indice_count=%ld", (long)od.sizes[0], (long)idx_desc.size);
  auto* comm        = h->comm;
  const size_t et   = wholememory_dtype_get_element_size(td.dtype);
  const size_t eo   = wholememory_dtype_get_element_size(od.dtype);
  const int64_t D   = td.sizes[1];
  const size_t isz  = idx_desc.dtype == WHOLEMEMORY_DT_INT64 ? 8 : 4;
  const char* idx_p = static_cast<const char*>(indices) + idx_desc.storage_offset * isz;

  exchange_plan p(env);
  plan_exchange(&p, comm, idx_p, idx_desc.dtype, idx_desc.size, handle_first_rows(h, et * td.stride), stream);

  /* owner side: gather requested rows (converted to the output dtype) into a packed buffer */
  temp_buffer served(env), fetched(env);
  void* served_p  = served.device((size_t)std::max<int64_t>(p.n_recv, 1) * D, od.dtype);
  void* fetched_p = fetched.device((size_t)std::max<int64_t>(p.n_send, 1) * D, od.dtype);
  int64_t sz[2]   = {p.n_recv, D};
  auto packed_d   = wholememory_create_matrix_desc(sz, D, 0, od.dtype);
  if (p.n_recv > 0)
    row_move(true, local_shard_ref(h), td, p.recv_idx.ptr(), wholememory_create_array_desc(p.n_recv, 0, idx_desc.dtype), served_p,
             packed_d, stream, sms);
  exchange_rows(p, comm, served_p, fetched_p, (size_t)D * eo, /*to_owner=*/false, stream);
  /* requester side: un-permute into the caller's output */
  if (p.n_send > 0) {
    int64_t fsz[2] = {p.n_send, D};
    auto fetched_d = wholememory_create_matrix_desc(fsz, D, 0, od.dtype);
    row_move(false, make_flat_table_ref(output), od, p.origin.ptr(), wholememory_create_array_desc(p.n_send, 0, WHOLEMEMORY_DT_INT64),
             fetched_p, fetched_d, stream, sms);
  }
  /* temp buffers are released when this frame unwinds: drain the stream first */
  WM_CUDA(cudaStreamSynchronize(stream));
}

void scatter_by_exchange(const void* input,
                         const wholememory_matrix_description_t& id,
                         const void* indices,
                         const wholememory_array_description_t& idx_desc,
                         wholememory_handle_t h,
                         const wholememory_matrix_description_t& td,
                         wholememory_env_func_t* env,
                         cudaStream_t stream,
                         int sms)
{
  require_cuda("wholememory_scatter (DISTRIBUTED)");
  WM_EXPECT(td.storage_offset >= 0 && td.storage_offset + td.sizes[1] <= td.stride, WHOLEMEMORY_INVALID_INPUT,
            "DISTRIBUTED scatter needs a tensor that starts at row 0");
  WM_EXPECT(id.sizes[0] == idx_desc.size, WHOLEMEMORY_LOGIC_ERROR, "input rows=%ld but indice_count=%ld", (long)id.sizes[0], (long)idx_desc.size);
  auto* comm        = h->comm;
  const size_t et   = wholememory_dtype_get_element_size(td.dtype);
  const size_t ei   = wholememory_dtype_get_element_size(id.dtype);
  const int64_t D   = td.sizes[1];
  const size_t isz  = idx_desc.dtype == WHOLEMEMORY_DT_INT64 ? 8 : 4;
  const char* idx_p = static_cast<const char*>(indices) + idx_desc.storage_offset * isz;

  exchange_plan p(env);
  plan_exchange(&p, comm, idx_p, idx_desc.dtype, idx_desc.size, handle_first_rows(h, et * td.stride), stream);

  temp_buffer outgoing(env), incoming(env);
  void* out_p   = outgoing.device((size_t)std::max<int64_t>(p.n_send, 1) * D, id.dtype);
  void* in_p    = incoming.device((size_t)std::max<int64_t>(p.n_recv, 1) * D, id.dtype);
  /* requester side: pack input rows in owner order */
  if (p.n_send > 0) {
    int64_t sz[2] = {p.n_send, D};
    auto packed_d = wholememory_create_matrix_desc(sz, D, 0, id.dtype);
    row_move(true, make_flat_table_ref(const_cast<void*>(input)), id, p.origin.ptr(),
             wholememory_create_array_desc(p.n_send, 0, WHOLEMEMORY_DT_INT64), out_p, packed_d, stream, sms);
  }
  exchange_rows(p, comm, out_p, in_p, (size_t)D * ei, /*to_owner=*/true, stream);
  /* owner side: store (and convert) into the local shard */
  if (p.n_recv > 0) {
    int64_t sz[2] = {p.n_recv, D};
    auto recv_d   = wholememory_create_matrix_desc(sz, D, 0, id.dtype);
    row_move(false, local_shard_ref(h), td, p.recv_idx.ptr(), wholememory_create_array_desc(p.n_recv, 0, idx_desc.dtype), in_p, recv_d,
             stream, sms);
  }
  WM_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace wm
