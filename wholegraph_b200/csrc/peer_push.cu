/*
 * Gradient exchange by PEER STORES (NVLink / NVSwitch) instead of an all-to-all.
 *
 * Replaces, for peer-mapped embeddings, the requester half of the reference's gradient pipeline
 * (cpp/src/wholememory/embedding.cpp:146-323: bucket + sort the ids, gather-permute the gradient rows into a send
 * buffer, exchange_ids / exchange_embeddings_nccl_func.cu:32-74 alltoallv of ids and of rows).
 *
 * Every rank owns one staging area inside a CHUNKED/DEVICE WholeMemory allocation that all ranks map.  After the ids
 * are grouped by owner (exchange.cu partition) and the world_size^2 count matrix is known on every rank (one
 * allgather over the bootstrap sockets), ONE kernel per rank reads its gradient rows in caller order and stores each
 * (id, row) straight into the owner's stage at its final position: no packed send buffer, no NCCL staging copies, each
 * gradient byte crosses NVLink exactly once as a store.  The owner then runs the fused merge + optimizer kernel
 * (sparse_optimizer.cu) on its stage.  Arrival layout is (sender rank, sender order) - the order the all-to-all
 * delivers - so duplicate-gradient sums are bit-identical to the NCCL path.
 *
 * Synchronisation: stores complete (cudaStreamSynchronize) -> bootstrap barrier -> owner kernels.  The stage is
 * double-buffered, so the owner's update kernel of step k can still be running while step k+1's rows arrive; a buffer
 * is overwritten in step k+2, which every rank enters only after its own step k+1 synchronisation.
 * Traffic per rank: read n*D*4 local, write n*(D*4+8) of which (ws-1)/ws over NVLink.
 */
#include "exchange.hpp"

#include <algorithm>

namespace wm {

namespace {

constexpr int kPushWarps = 8;

struct push_dest {
  int nranks;
  int64_t first_logical; /* grouped position the grid starts at: bucket (me + 1) % ws, see push_rows_to_owners */
  int64_t bucket_start[kMaxInlineRanks + 1]; /* grouped order: bucket r is [bucket_start[r], bucket_start[r+1]) */
  int64_t dest_row[kMaxInlineRanks];         /* first row of MY block inside owner r's stage */
  int64_t* ids[kMaxInlineRanks];             /* owner r's id array (mapped here) */
  float* rows[kMaxInlineRanks];              /* owner r's row array */
};

template <typename IdxT, int VEC>
__global__ void __launch_bounds__(kPushWarps * 32) push_rows_kernel(const IdxT* __restrict__ grouped_idx,
                                                                    const int64_t* __restrict__ origin,
                                                                    int64_t n_send,
                                                                    const float* __restrict__ rows_in,
                                                                    int64_t row_stride,
                                                                    int dim,
                                                                    push_dest d)
{
  const int lane = threadIdx.x & 31;
  int64_t j      = (int64_t)blockIdx.x * kPushWarps + (threadIdx.x >> 5);
  if (j >= n_send) return;
  /* the grid walks the owner buckets ROTATED: rank `me` starts with owner me + 1 and ends with itself, so at any moment the
   * ranks of the box store into different owners instead of all flooding owner 0, then owner 1, ... (one ingress link busy,
   * the others idle: measured 3.05 ms for the whole step at 8 GPUs, slower than the NCCL all-to-all's 2.43 ms) */
  j += d.first_logical;
  if (j >= n_send) j -= n_send;
  const int64_t src_row = origin[j];
  const IdxT id         = grouped_idx[j];
  int r = 0;
#pragma unroll 1
  for (int q = 1; q < d.nranks; ++q)
    if (j >= d.bucket_start[q]) r = q;
  const int64_t dst_row = d.dest_row[r] + (j - d.bucket_start[r]);
  if (lane == 0) d.ids[r][dst_row] = (int64_t)id;
  const float* src = rows_in + src_row * row_stride;
  float* dst       = d.rows[r] + dst_row * dim;
  if constexpr (VEC == 4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4       = reinterpret_cast<float4*>(dst);
    const int nv     = dim >> 2;
    for (int c0 = lane; c0 < nv; c0 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u * 32 < nv) v[u] = s4[c0 + u * 32];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + u * 32 < nv) d4[c0 + u * 32] = v[u];
    }
  } else {
    for (int c = lane; c < dim; c += 32) dst[c] = src[c];
  }
}

size_t stage_bytes_per_rank(int64_t cap_rows, int64_t dim) { return 2 * (size_t)cap_rows * ((size_t)dim * sizeof(float) + sizeof(int64_t)); }

/* buffer b of rank r: ids then rows */
int64_t* stage_ids(const push_stage& st, int r, int b)
{
  char* base = static_cast<char*>(st.h->rank_base[r]) + (size_t)b * (stage_bytes_per_rank(st.cap_rows, st.dim) / 2);
  return reinterpret_cast<int64_t*>(base);
}
float* stage_rows(const push_stage& st, int r, int b)
{
  return reinterpret_cast<float*>(reinterpret_cast<char*>(stage_ids(st, r, b)) + (size_t)st.cap_rows * sizeof(int64_t));
}

}  // namespace

void destroy_push_stage(push_stage* st)
{
  if (st->h != nullptr) {
    wholememory_free(st->h);
    st->h = nullptr;
  }
  st->cap_rows = 0;
}

int64_t push_rows_to_owners(push_stage* st,
                            wholememory_comm_t comm,
                            const exchange_plan& p,
                            const float* rows_in,
                            int64_t row_stride,
                            int64_t dim,
                            cudaStream_t stream,
                            const int64_t** ids,
                            const float** rows)
{
  const int ws = comm->world_size, me = comm->world_rank;
  WM_EXPECT(ws <= kMaxInlineRanks, WHOLEMEMORY_NOT_SUPPORTED, "peer push supports at most %d ranks", kMaxInlineRanks);
  /* count matrix C[q][r] = rows rank q sends to rank r */
  std::vector<int64_t> counts((size_t)ws * ws);
  {
    std::lock_guard<std::mutex> lk(comm->mu);
    comm->boot->allgather(p.send_counts.data(), counts.data(), sizeof(int64_t) * ws);
  }
  int64_t max_need = 0, n_recv = 0;
  for (int r = 0; r < ws; ++r) {
    int64_t need = 0;
    for (int q = 0; q < ws; ++q) need += counts[(size_t)q * ws + r];
    max_need = std::max(max_need, need);
    if (r == me) n_recv = need;
  }
  /* every rank sees the same matrix, so every rank takes the same (collective) decision to grow */
  if (st->h == nullptr || max_need > st->cap_rows || dim != st->dim) {
    WM_CUDA(cudaStreamSynchronize(stream));
    destroy_push_stage(st);
    int64_t cap = std::max<int64_t>(4096, max_need + max_need / 4);
    cap         = (cap + 1023) / 1024 * 1024; /* even => the row array stays 16-byte aligned */
    st->dim     = dim;
    const size_t per_rank = stage_bytes_per_rank(cap, dim);
    wholememory_error_code_t rc =
      wholememory_malloc(&st->h, per_rank * (size_t)ws, comm, WHOLEMEMORY_MT_CHUNKED, WHOLEMEMORY_ML_DEVICE, per_rank, nullptr);
    WM_EXPECT(rc == WHOLEMEMORY_SUCCESS && st->h != nullptr && st->h->peer_mapped, rc == WHOLEMEMORY_SUCCESS ? WHOLEMEMORY_LOGIC_ERROR : rc,
              "cannot allocate the %zu-byte gradient stage", per_rank * (size_t)ws);
    st->cap_rows = cap;
    st->flip     = 0;
  }
  const int b = st->flip;
  st->flip ^= 1;

  if (p.n_send > 0) {
    push_dest d{};
    d.nranks    = ws;
    int64_t acc = 0;
    for (int r = 0; r < ws; ++r) {
      d.bucket_start[r] = acc;
      acc += p.send_counts[r];
      int64_t before = 0;
      for (int q = 0; q < me; ++q) before += counts[(size_t)q * ws + r];
      d.dest_row[r] = before;
      d.ids[r]      = stage_ids(*st, r, b);
      d.rows[r]     = stage_rows(*st, r, b);
    }
    d.bucket_start[ws] = acc;
    d.first_logical    = d.bucket_start[(me + 1) % ws];
    const bool vec4 = dim % 4 == 0 && row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(rows_in) & 15) == 0;
    const unsigned grid = (unsigned)((p.n_send + kPushWarps - 1) / kPushWarps);
    const bool idx64    = p.idx_dtype == WHOLEMEMORY_DT_INT64;
    auto* org           = static_cast<const int64_t*>(p.origin.ptr());
    if (idx64 && vec4)
      push_rows_kernel<int64_t, 4><<<grid, kPushWarps * 32, 0, stream>>>(static_cast<const int64_t*>(p.grouped_idx.ptr()), org, p.n_send, rows_in, row_stride, (int)dim, d);
    else if (idx64)
      push_rows_kernel<int64_t, 1><<<grid, kPushWarps * 32, 0, stream>>>(static_cast<const int64_t*>(p.grouped_idx.ptr()), org, p.n_send, rows_in, row_stride, (int)dim, d);
    else if (vec4)
      push_rows_kernel<int32_t, 4><<<grid, kPushWarps * 32, 0, stream>>>(static_cast<const int32_t*>(p.grouped_idx.ptr()), org, p.n_send, rows_in, row_stride, (int)dim, d);
    else
      push_rows_kernel<int32_t, 1><<<grid, kPushWarps * 32, 0, stream>>>(static_cast<const int32_t*>(p.grouped_idx.ptr()), org, p.n_send, rows_in, row_stride, (int)dim, d);
    WM_CUDA(cudaGetLastError());
  }
  /* my stores have landed at their owners once the stream drains; everyone's have after the barrier */
  WM_CUDA(cudaStreamSynchronize(stream));
  {
    std::lock_guard<std::mutex> lk(comm->mu);
    comm->boot->barrier();
  }
  *ids  = stage_ids(*st, me, b);
  *rows = stage_rows(*st, me, b);
  return n_recv;
}

}  // namespace wm
