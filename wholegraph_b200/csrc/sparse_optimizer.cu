/*
 * Duplicate-gradient merge FUSED with the sparse optimizer row update (one kernel, one pass over
 * each touched row of W and its state).
 *
 * Replaces, as ONE kernel, the reference's three stages:
 *   dedup_indice_and_gradients   functions/exchange_embeddings_nccl_func.cu:76-206
 *        (sort + thrust::unique_by_key + DedupIndiceAndGradientsKernel writing a deduped grad matrix)
 *   {sgd,lazy_adam,ada_grad,rms_prop}_optimizer_step_kernel
 *        functions/embedding_optimizer_func.cu:178-224, :331-419, :594-657, :791-851
 * Arithmetic (expression order, per-row beta^t handling, sequential left-to-right summation of
 * duplicate gradients in arrival order) is kept identical so results match the reference's.
 *
 * Design: received (row id, position) pairs are radix-sorted by row id (stable => arrival order
 * inside a run).  The kernel is launched with one WARP per SORTED POSITION (8 per CTA); a warp whose
 * position is not the head of a run exits at once, a head warp issues every load of its row (gradient,
 * W, state) before the first use, walks its run summing duplicate gradient rows in registers, and then
 * applies the optimizer to W / state in place.  No unique-count is needed on the host (no D2H sync),
 * no deduplicated gradient matrix is written or re-read, and the row of W is touched by exactly one
 * kernel.  Rows are moved as float4 when alignment allows.
 * Roofline: per unique row  read g*dups + W + state, write W + state  (7*D*4+16 B for LazyAdam,
 * D=512 -> 14,352 B), all local HBM.
 */
#include "sparse_optimizer.hpp"

#include <cub/device/device_radix_sort.cuh>

namespace wm {

namespace {

template <int VEC>
struct fvec;
template <>
struct fvec<4> {
  float4 v;
  __device__ __forceinline__ float& at(int i) { return (&v.x)[i]; }
};
template <>
struct fvec<1> {
  float v;
  __device__ __forceinline__ float& at(int) { return v; }
};

template <int VEC>
__device__ __forceinline__ fvec<VEC> ldv(const float* p)
{
  fvec<VEC> r;
  if constexpr (VEC == 4) r.v = *reinterpret_cast<const float4*>(p);
  else r.v = *p;
  return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, fvec<VEC> x)
{
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = x.v;
  else *p = x.v;
}

__global__ void iota_kernel(int* p, int n, int* long_count)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *long_count = 0;
  if (i < n) p[i] = i;
}

/* One optimizer step on VEC consecutive elements of a row.  Expression order follows the reference kernels
 * (embedding_optimizer_func.cu:178-224 SGD, :331-419 LazyAdam, :594-657 AdaGrad, :791-851 RMSProp) so results match bit for bit. */
template <int OPT, int VEC>
__device__ __forceinline__ void optimizer_step(fvec<VEC>& g, fvec<VEC>& wv, fvec<VEC>& sv0, fvec<VEC>& sv1, const optimizer_params& p,
                                               float lr, float beta1t, float beta2t)
{
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float grad_value      = g.at(k);
    float embedding_value = wv.at(k);
    if (OPT == WHOLEMEMORY_OPT_SGD) {
      grad_value += p.weight_decay * embedding_value;
      embedding_value -= lr * grad_value;
    } else if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) {
      if (p.adam_w) {
        embedding_value -= lr * p.weight_decay * embedding_value;
      } else {
        grad_value = grad_value + p.weight_decay * embedding_value;
      }
      float m         = sv0.at(k);
      float v         = sv1.at(k);
      m               = p.beta1 * m + (1 - p.beta1) * grad_value;
      v               = p.beta2 * v + (1 - p.beta2) * grad_value * grad_value;
      float mhat      = m / (1 - beta1t);
      float vhat      = v / (1 - beta2t);
      embedding_value = embedding_value - lr * mhat / (sqrtf(vhat) + p.epsilon);
      sv0.at(k)       = m;
      sv1.at(k)       = v;
    } else if (OPT == WHOLEMEMORY_OPT_ADAGRAD) {
      grad_value      = grad_value + p.weight_decay * embedding_value;
      float state_sum = sv0.at(k);
      state_sum       = state_sum + grad_value * grad_value;
      embedding_value = embedding_value - lr * grad_value / (sqrtf(state_sum) + p.epsilon);
      sv0.at(k)       = state_sum;
    } else { /* RMSPROP */
      grad_value      = grad_value + p.weight_decay * embedding_value;
      float v         = sv0.at(k);
      v               = p.alpha * v + (1 - p.alpha) * grad_value * grad_value;
      embedding_value = embedding_value - lr * grad_value / (sqrtf(v) + p.epsilon);
      sv0.at(k)       = v;
    }
    wv.at(k) = embedding_value;
  }
}

constexpr int kLongRun     = 64;  /* runs longer than this are merged by long_run_update_kernel (a CTA per run) */
constexpr int kLongThreads = 128;
constexpr int kLongDepth   = 8;   /* gradient rows in flight per thread, times two buffers */
constexpr int kOptWarps   = 4; /* sorted positions per CTA (one warp each) */
constexpr int kOptMinCtas = 8; /* <= 64 registers: 32 resident warps per SM */

/* One WARP per sorted position.  Each lane keeps U vectors of every stream (gradient, W, state) in registers, so
 * U*4 independent 16 B loads per lane are issued before the first use.  Measured on B200 (LazyAdam, 512-float rows,
 * tools/bench_ops.py, whole call): CTA per position 0.794 ms; warp per position U=4 (96 regs) 0.897, U=2 8-warp CTAs
 * 0.784, U=1 0.778; U=2 with 4-warp CTAs capped at 64 registers 0.710 <- this configuration.  The step has ~55
 * instructions per element (three IEEE divides and a square root), so resident warps matter more than loads per lane. */
/* positions 0..n-1 plus a copy of the ids with every id outside [0, total_rows) replaced by total_rows */
template <typename IdxT>
__global__ void iota_fold_kernel(const IdxT* __restrict__ ids, IdxT* __restrict__ folded, int* __restrict__ pos, int n, IdxT total_rows,
                                 int* __restrict__ long_count)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *long_count = 0; /* the list of long-run heads starts empty (no separate memset launch) */
  if (i >= n) return;
  const IdxT v = ids[i];
  folded[i]    = (v < 0 || v >= total_rows) ? total_rows : v;
  pos[i]       = i;
}

template <typename IdxT, int OPT, int VEC, int U>
__global__ void __launch_bounds__(kOptWarps * 32, kOptMinCtas) fused_merge_update_kernel(const IdxT* __restrict__ sorted_idx,
                                                                            const int* __restrict__ sorted_pos,
                                                                            int n,
                                                                            const float* __restrict__ grads,
                                                                            int64_t grad_stride,
                                                                            optimizer_rows rows,
                                                                            optimizer_params p,
                                                                            float lr,
                                                                            int* __restrict__ long_heads /* [0] = count, then head positions */)
{
  const int lane = threadIdx.x & 31;
  const int b    = blockIdx.x * kOptWarps + (threadIdx.x >> 5);
  if (b >= n) return;
  /* independent loads up front (one DRAM round trip): my id, my neighbours' ids, my gradient position */
  const IdxT row_id  = sorted_idx[b];
  const IdxT prev_id = b > 0 ? sorted_idx[b - 1] : row_id;
  const IdxT next_id = b + 1 < n ? sorted_idx[b + 1] : row_id;
  const IdxT far_id  = b + kLongRun < n ? sorted_idx[b + kLongRun] : row_id;
  const int pos0     = sorted_pos[b];
  if (b > 0 && prev_id == row_id) return; /* not the head of its run */
  const bool has_dups = b + 1 < n && next_id == row_id;
  if (b + kLongRun < n && far_id == row_id) { /* more than kLongRun gradients for this row: leave it to long_run_update_kernel */
    if (lane == 0) long_heads[1 + atomicAdd(long_heads, 1)] = b;
    return;
  }
  const int64_t local = (int64_t)row_id - rows.local_row_start;
  if (local < 0 || local >= rows.local_rows) return; /* negative / foreign ids are ignored */

  float* w  = rows.w + local * rows.w_stride;
  float* s0 = rows.state ? rows.state + local * rows.state_stride : nullptr; /* m | state_sum | v */
  float* s1 = s0 ? s0 + rows.w_stride : nullptr;                             /* LazyAdam v */
  const float* g0 = grads + (int64_t)pos0 * grad_stride;

  float beta1t = 0.f, beta2t = 0.f;
  if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) {
    /* per-row powers are advanced BEFORE use (embedding_optimizer_func.cu:386-389) */
    beta1t = rows.b12[local * 2 + 0] * p.beta1;
    beta2t = rows.b12[local * 2 + 1] * p.beta2;
  }

  constexpr int kStep = 32 * VEC;
  for (int c0 = lane * VEC; c0 < rows.dim; c0 += kStep * U) {
    fvec<VEC> g[U], wv[U], sv0[U], sv1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * kStep;
      if (c < rows.dim) {
        g[u]  = ldv<VEC>(g0 + c);
        wv[u] = ldv<VEC>(w + c);
        if (OPT != WHOLEMEMORY_OPT_SGD) sv0[u] = ldv<VEC>(s0 + c);
        if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) sv1[u] = ldv<VEC>(s1 + c);
      }
    }
    /* 1. merge duplicates: g = g[pos0] + g[pos1] + ... in arrival order */
    if (has_dups) {
      for (int j = b + 1; j < n && sorted_idx[j] == row_id; ++j) {
        const float* gj = grads + (int64_t)sorted_pos[j] * grad_stride;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + u * kStep;
          if (c < rows.dim) {
            fvec<VEC> o = ldv<VEC>(gj + c);
#pragma unroll
            for (int k = 0; k < VEC; ++k) g[u].at(k) += o.at(k);
          }
        }
      }
    }
    /* 2. optimizer, in place */
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + u * kStep;
      if (c < rows.dim) {
        optimizer_step<OPT, VEC>(g[u], wv[u], sv0[u], sv1[u], p, lr, beta1t, beta2t);
        if (OPT != WHOLEMEMORY_OPT_SGD) stv<VEC>(s0 + c, sv0[u]);
        if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) stv<VEC>(s1 + c, sv1[u]);
        stv<VEC>(w + c, wv[u]);
      }
    }
  }
  if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM && lane == 0) {
    rows.b12[local * 2 + 0] = beta1t;
    rows.b12[local * 2 + 1] = beta2t;
  }
}


/*
 * Hot rows.  A row that receives thousands of gradients in one step (hub nodes of a power-law graph: with Zipf(1.05) ids
 * one row collects ~5 % of all gradients) is a serial chain for the warp-per-position kernel above -- one dependent
 * (position, row) load pair per duplicate, ~1.5 us each: 140 ms per step at 8 GPUs, and the reference's block-per-id
 * dedup kernel has the same shape (111 ms).  Runs longer than kLongRun are therefore left to this kernel: ONE CTA per run,
 * every thread owns one vector column of the row and walks the whole run IN ARRIVAL ORDER (the sum stays bit-identical to
 * the sequential one) with 2 x kLongDepth independent row loads in flight, then applies the optimizer to its column.
 * The heads of such runs are appended to a small device list by the warp-per-position kernel (no host round trip); this
 * kernel is a fixed grid whose CTAs take list entries round-robin and exit at once when the list is empty.
 */
template <typename IdxT, int OPT, int VEC>
__global__ void __launch_bounds__(kLongThreads) long_run_update_kernel(const IdxT* __restrict__ sorted_idx,
                                                                      const int* __restrict__ sorted_pos,
                                                                      int n,
                                                                      const float* __restrict__ grads,
                                                                      int64_t grad_stride,
                                                                      optimizer_rows rows,
                                                                      optimizer_params p,
                                                                      float lr,
                                                                      const int* __restrict__ long_heads)
{
  const int t     = threadIdx.x;
  const int nlong = long_heads[0];
  __shared__ int s_end;
  for (int item = blockIdx.x; item < nlong; item += gridDim.x) {
    const int head    = long_heads[1 + item];
    const IdxT row_id = sorted_idx[head];
    /* end: first position whose id differs (binary search, ids ascending; the run is known to reach head + kLongRun) */
    __syncthreads(); /* s_end of the previous item has been read by everyone */
    if (t == 0) {
      int lo = head + kLongRun, hi = n; /* invariant: sorted_idx[lo] == row_id, (hi == n or sorted_idx[hi] > row_id) */
      while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        if (sorted_idx[mid] == row_id) lo = mid;
        else hi = mid;
      }
      s_end = hi;
    }
    __syncthreads();
    const int end = s_end;
  const int64_t local = (int64_t)row_id - rows.local_row_start;
  if (local < 0 || local >= rows.local_rows) continue; /* the fold sentinel / foreign ids (uniform over the CTA) */

  float* w  = rows.w + local * rows.w_stride;
  float* s0 = rows.state ? rows.state + local * rows.state_stride : nullptr;
  float* s1 = s0 ? s0 + rows.w_stride : nullptr;
  float beta1t = 0.f, beta2t = 0.f;
  if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) {
    beta1t = rows.b12[local * 2 + 0] * p.beta1;
    beta2t = rows.b12[local * 2 + 1] * p.beta2;
  }
  for (int c = t * VEC; c < rows.dim; c += kLongThreads * VEC) {
    fvec<VEC> acc = ldv<VEC>(grads + (int64_t)sorted_pos[head] * grad_stride + c);
    fvec<VEC> wv  = ldv<VEC>(w + c), sv0, sv1;
    if (OPT != WHOLEMEMORY_OPT_SGD) sv0 = ldv<VEC>(s0 + c);
    if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) sv1 = ldv<VEC>(s1 + c);
    fvec<VEC> a[kLongDepth], b[kLongDepth];
    auto fetch = [&](int j0, fvec<VEC>* buf) {
#pragma unroll
      for (int k = 0; k < kLongDepth; ++k)
        if (j0 + k < end) buf[k] = ldv<VEC>(grads + (int64_t)sorted_pos[j0 + k] * grad_stride + c);
    };
    auto add = [&](int j0, fvec<VEC>* buf) { /* strictly in arrival order */
#pragma unroll
      for (int k = 0; k < kLongDepth; ++k)
        if (j0 + k < end) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc.at(e) += buf[k].at(e);
        }
    };
    int j = head + 1;
    fetch(j, a);
    while (j < end) {
      fetch(j + kLongDepth, b);
      add(j, a);
      j += kLongDepth;
      if (j >= end) break;
      fetch(j + kLongDepth, a);
      add(j, b);
      j += kLongDepth;
    }
    optimizer_step<OPT, VEC>(acc, wv, sv0, sv1, p, lr, beta1t, beta2t);
    if (OPT != WHOLEMEMORY_OPT_SGD) stv<VEC>(s0 + c, sv0);
    if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) stv<VEC>(s1 + c, sv1);
    stv<VEC>(w + c, wv);
  }
  if (OPT == WHOLEMEMORY_OPT_LAZY_ADAM) {
    __syncthreads(); /* every thread has read the old powers */
    if (t == 0) {
      rows.b12[local * 2 + 0] = beta1t;
      rows.b12[local * 2 + 1] = beta2t;
    }
  }
  } /* next list entry */
}

template <typename IdxT, int VEC>
void launch_long_runs(int opt, const IdxT* si, const int* sp, int n, const float* g, int64_t gs, const optimizer_rows& rows,
                      const optimizer_params& p, float lr, const int* long_heads, cudaStream_t s)
{
  if (n <= kLongRun) return; /* no run can be longer than kLongRun: the list is empty */
  const unsigned grid = (unsigned)std::min<int64_t>((n + kLongRun - 1) / kLongRun, 2 * (int64_t)sm_count());
  switch (opt) {
    case WHOLEMEMORY_OPT_SGD: long_run_update_kernel<IdxT, WHOLEMEMORY_OPT_SGD, VEC><<<grid, kLongThreads, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads); break;
    case WHOLEMEMORY_OPT_LAZY_ADAM: long_run_update_kernel<IdxT, WHOLEMEMORY_OPT_LAZY_ADAM, VEC><<<grid, kLongThreads, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads); break;
    case WHOLEMEMORY_OPT_ADAGRAD: long_run_update_kernel<IdxT, WHOLEMEMORY_OPT_ADAGRAD, VEC><<<grid, kLongThreads, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads); break;
    case WHOLEMEMORY_OPT_RMSPROP: long_run_update_kernel<IdxT, WHOLEMEMORY_OPT_RMSPROP, VEC><<<grid, kLongThreads, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads); break;
    default: break;
  }
}

template <typename IdxT, int VEC, int U>
void launch_fused_u(int opt, const IdxT* si, const int* sp, int n, const float* g, int64_t gs, const optimizer_rows& rows,
                    const optimizer_params& p, float lr, int* long_heads, cudaStream_t s)
{
  const unsigned grid = (unsigned)((n + kOptWarps - 1) / kOptWarps);
  const unsigned cta  = kOptWarps * 32;
  switch (opt) {
    case WHOLEMEMORY_OPT_SGD:
      fused_merge_update_kernel<IdxT, WHOLEMEMORY_OPT_SGD, VEC, U><<<grid, cta, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads);
      break;
    case WHOLEMEMORY_OPT_LAZY_ADAM:
      fused_merge_update_kernel<IdxT, WHOLEMEMORY_OPT_LAZY_ADAM, VEC, U><<<grid, cta, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads);
      break;
    case WHOLEMEMORY_OPT_ADAGRAD:
      fused_merge_update_kernel<IdxT, WHOLEMEMORY_OPT_ADAGRAD, VEC, U><<<grid, cta, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads);
      break;
    case WHOLEMEMORY_OPT_RMSPROP:
      fused_merge_update_kernel<IdxT, WHOLEMEMORY_OPT_RMSPROP, VEC, U><<<grid, cta, 0, s>>>(si, sp, n, g, gs, rows, p, lr, long_heads);
      break;
    default: WM_THROW(WHOLEMEMORY_INVALID_INPUT, "unknown optimizer type %d", opt);
  }
}

template <typename IdxT, int VEC>
void launch_fused(int opt, const IdxT* si, const int* sp, int n, const float* g, int64_t gs, const optimizer_rows& rows,
                  const optimizer_params& p, float lr, int* long_heads, cudaStream_t s)
{
  if (rows.dim > 32 * VEC) launch_fused_u<IdxT, VEC, 2>(opt, si, sp, n, g, gs, rows, p, lr, long_heads, s);
  else launch_fused_u<IdxT, VEC, 1>(opt, si, sp, n, g, gs, rows, p, lr, long_heads, s);
}

template <typename IdxT>
void merge_update_typed(int opt, const void* idx, int64_t n, const float* grads, int64_t grad_stride, const optimizer_rows& rows,
                        const optimizer_params& p, float lr, int64_t total_rows, bool may_have_negative, wholememory_env_func_t* env,
                        cudaStream_t s)
{
  const wholememory_dtype_t idt = sizeof(IdxT) == 8 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT;
  temp_buffer sorted_idx_b(env), pos_in_b(env), pos_out_b(env), cub_b(env);
  auto* sorted_idx = static_cast<IdxT*>(sorted_idx_b.device((size_t)n, idt));
  auto* pos_in     = static_cast<int*>(pos_in_b.device((size_t)n, WHOLEMEMORY_DT_INT));
  auto* pos_out    = static_cast<int*>(pos_out_b.device((size_t)n, WHOLEMEMORY_DT_INT));
  temp_buffer long_b(env); /* [0] = number of runs longer than kLongRun, then their head positions (at most n / kLongRun) */
  auto* long_heads = static_cast<int*>(long_b.device((size_t)(n / kLongRun + 2), WHOLEMEMORY_DT_INT));
  /* Only the bits in use are sorted: ids below total_rows need ceil(log2(total_rows + 1)) bits, i.e. 3 onesweep passes
   * instead of 8 for 5M rows.  Caller-supplied ids may be negative or out of range (ignored by the update kernel) and
   * would alias into that range, so they are first folded onto ONE sentinel key, total_rows, in the pass that writes
   * the positions anyway. */
  const IdxT* keys_in = static_cast<const IdxT*>(idx);
  temp_buffer folded_b(env);
  int end_bit = (int)sizeof(IdxT) * 8;
  const bool fits = total_rows > 0 && (sizeof(IdxT) == 8 || total_rows < (int64_t)0x7fffffff);
  if (fits) {
    end_bit = 1;
    while (end_bit < (int)sizeof(IdxT) * 8 && ((int64_t)1 << end_bit) <= total_rows) ++end_bit;
  }
  if (fits && may_have_negative) {
    auto* folded = static_cast<IdxT*>(folded_b.device((size_t)n, idt));
    iota_fold_kernel<IdxT><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys_in, folded, pos_in, (int)n, (IdxT)total_rows, long_heads);
    keys_in = folded;
  } else {
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pos_in, (int)n, long_heads);
  }
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys_in, sorted_idx, pos_in, pos_out, (int)n, 0, end_bit, s);
  void* cub_tmp = cub_b.device(cub_bytes, WHOLEMEMORY_DT_INT8);
  cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, sorted_idx, pos_in, pos_out, (int)n, 0, end_bit, s);
  bool vec4 = rows.dim % 4 == 0 && grad_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0 &&
              rows.w_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(rows.w) & 15) == 0 &&
              (rows.state == nullptr || (rows.state_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(rows.state) & 15) == 0));
  if (vec4) launch_fused<IdxT, 4>(opt, sorted_idx, pos_out, (int)n, grads, grad_stride, rows, p, lr, long_heads, s);
  else launch_fused<IdxT, 1>(opt, sorted_idx, pos_out, (int)n, grads, grad_stride, rows, p, lr, long_heads, s);
  /* rows with more than kLongRun gradients (listed by the kernel above): one CTA per run, deep prefetch, same summation order */
  if (vec4) launch_long_runs<IdxT, 4>(opt, sorted_idx, pos_out, (int)n, grads, grad_stride, rows, p, lr, long_heads, s);
  else launch_long_runs<IdxT, 1>(opt, sorted_idx, pos_out, (int)n, grads, grad_stride, rows, p, lr, long_heads, s);
  WM_CUDA(cudaGetLastError());
  /* Temporaries go back to the caller's allocator when this frame unwinds WITHOUT a host sync, like the reference's
   * dedup/optimizer stage: torch's caching allocator is stream-ordered, cudaFree (default env) synchronises itself. */
}

__global__ void fill_kernel(float* p, float v, int64_t n)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

void merge_and_update_rows(int optimizer_type,
                           const void* indices,
                           wholememory_dtype_t idx_dtype,
                           int64_t n,
                           const float* grads,
                           int64_t grad_stride,
                           const optimizer_rows& rows,
                           const optimizer_params& params,
                           float lr,
                           int64_t total_rows,
                           bool may_have_negative,
                           wholememory_env_func_t* env,
                           cudaStream_t stream)
{
  require_cuda("sparse optimizer step");
  if (n == 0) return;
  WM_EXPECT(n < ((int64_t)1 << 31), WHOLEMEMORY_INVALID_VALUE, "too many gradient rows in one step (%ld)", (long)n);
  if (idx_dtype == WHOLEMEMORY_DT_INT64)
    merge_update_typed<int64_t>(optimizer_type, indices, n, grads, grad_stride, rows, params, lr, total_rows, may_have_negative, env, stream);
  else if (idx_dtype == WHOLEMEMORY_DT_INT)
    merge_update_typed<int32_t>(optimizer_type, indices, n, grads, grad_stride, rows, params, lr, total_rows, may_have_negative, env, stream);
  else
    WM_THROW(WHOLEMEMORY_LOGIC_ERROR, "gradient indices must be int32 or int64");
}

void fill_float(float* p, float value, int64_t n, cudaStream_t stream)
{
  if (n == 0) return;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, value, n);
  WM_CUDA(cudaGetLastError());
}

}  // namespace wm
