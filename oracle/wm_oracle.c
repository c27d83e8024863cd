/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product: nothing under wholegraph_b200/
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the CHECKER (or the timed CPU baseline).
 *
 * Plain-C restatement of the reference's hot path (rapidsai/wholegraph @ v24.12), one function
 * per reference kernel, each citing the file:line it follows (paths relative to the reference root).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   gather / scatter / partition : pinned -- against the reference tests' closed-form table pattern
 *       (cpp/tests/wholememory_ops/embedding_test_utils.cu:197-238) in tests/test_oracle.py, against an
 *       independent numpy restatement, and against the reference's own GPU kernels rebuilt from
 *       /root/reference (oracle/_ref/libwholegraph_ref.so) on the GPU box (tests/test_ref_parity_gpu.py) and their committed
 *       golden outputs (tests/golden/reference_gather_scatter_golden.npz, tests/test_golden.py).
 *   sparse optimizers            : restated from the kernels; pinned on CPU against the reference's own CPU test model
 *       compiled as code (class CPUOptimizer, cpp/tests/wholememory_ops/wholememory_embedding_gradient_apply_tests.cu:169-371,
 *       built by oracle/build_ref_host_optimizer_model.sh): bit-identical over multi-step schedules with duplicate ids
 *       (tests/test_ref_optimizer_model.py); the asserted tolerance stays the reference's own 1e-5 (:481-501).  The
 *       reference's optimizer kernels also build into oracle/_ref (RAFT stubbed by a declaration): the three-way GPU
 *       comparison (tests/test_zz_ref_optimizer_parity_gpu.py) is green on a B200 and their outputs are committed as
 *       tests/golden/reference_optimizer_golden.npz, which tests/test_golden_more.py holds this oracle to on CPU.
 *   neighbor sampler             : the selection algorithms are pinned against the reference's CPU models compiled as
 *       code (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:419-505 and :676-763, built by
 *       oracle/build_ref_host_sampling_model.sh; tests/test_ref_sampling_model.py, element for element) and against the
 *       reference's sampler KERNELS run on a restated generator (tests/test_zz_ref_sampling_parity_gpu.py, golden outputs in
 *       tests/golden/reference_sampler_golden.npz).  The RANDOM
 *       STREAM (RAFT PCGenerator, un-vendored dependency rapidsai/raft branch-24.12) is restated from the published
 *       PCG-XSH-RR 64/32 algorithm and checked against the pcg32 known-answer vector, not against RAFT itself:
 *       "parity unpinned" for the stream (no RAFT source / golden here).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* same numeric values as include/wholememory/tensor_description.h (reference tensor_description.h:29-40) */
enum { DT_UNKNOWN = 0, DT_FLOAT, DT_HALF, DT_DOUBLE, DT_BF16, DT_INT, DT_INT64, DT_INT16, DT_INT8 };

static int dt_size(int dt)
{
  switch (dt) {
    case DT_INT8: return 1;
    case DT_INT16:
    case DT_BF16:
    case DT_HALF: return 2;
    case DT_INT:
    case DT_FLOAT: return 4;
    case DT_INT64:
    case DT_DOUBLE: return 8;
    default: return 0;
  }
}
static int dt_is_float(int dt) { return dt == DT_FLOAT || dt == DT_HALF || dt == DT_DOUBLE || dt == DT_BF16; }

/* ---- IEEE binary16 / bfloat16 <-> binary32, round-to-nearest-even (what CUDA's
 * static_cast<__half>(float) / __float2bfloat16_rn do; type_caster, gather_scatter_func.cuh:161-208) ---- */
static float half_to_float(uint16_t h)
{
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp  = (h >> 10) & 0x1fu;
  uint32_t man  = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while ((man & 0x400u) == 0);
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

static uint16_t float_to_half(float f)
{
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t abs  = x & 0x7fffffffu;
  if (abs > 0x7f800000u) return (uint16_t)(sign | 0x7fffu); /* NaN -> canonical 0x7fff like cvt.rn.f16.f32 */
  if (abs >= 0x477ff000u) {                                 /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7c00u);
  }
  if (abs < 0x33000001u) return (uint16_t)sign; /* < 2^-25 (or exactly 2^-25, ties-to-even) -> 0 */
  int32_t exp  = (int32_t)(abs >> 23) - 127;
  uint32_t man = (abs & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t hexp;
  if (exp < -14) { /* subnormal half */
    shift = 13 + (-14 - exp);
    hexp  = 0;
  } else {
    shift = 13;
    hexp  = (uint32_t)(exp + 15);
  }
  uint32_t q    = man >> shift;
  uint32_t rem  = man & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  uint32_t r;
  if (hexp == 0)
    r = q; /* q may carry into exponent 1: correct by construction */
  else
    r = ((hexp - 1) << 10) + q; /* q includes the implicit bit (0x400), carry propagates */
  return (uint16_t)(sign | r);
}

static float bf16_to_float(uint16_t b)
{
  uint32_t bits = (uint32_t)b << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static uint16_t float_to_bf16(float f)
{
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return 0x7fffu; /* NaN, as __float2bfloat16_rn */
  uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (uint16_t)(x >> 16);
}

/* Load element as the reference's "LoadTypeT" (half/bf16 -> float, others themselves), widened to a
 * common carrier: floats travel as double-or-float tagged value, ints as int64. */
typedef struct {
  int is_f32; /* value passed through float (half/bf16/float sources) */
  float f;
  double d;
  int64_t i;
} elem_t;

static elem_t load_elem(const void* p, int dt)
{
  elem_t e;
  memset(&e, 0, sizeof(e));
  switch (dt) {
    case DT_FLOAT: e.is_f32 = 1; e.f = *(const float*)p; break;
    case DT_HALF: e.is_f32 = 1; e.f = half_to_float(*(const uint16_t*)p); break;
    case DT_BF16: e.is_f32 = 1; e.f = bf16_to_float(*(const uint16_t*)p); break;
    case DT_DOUBLE: e.d = *(const double*)p; break;
    case DT_INT8: e.i = *(const int8_t*)p; break;
    case DT_INT16: e.i = *(const int16_t*)p; break;
    case DT_INT: e.i = *(const int32_t*)p; break;
    case DT_INT64: e.i = *(const int64_t*)p; break;
    default: break;
  }
  return e;
}

/* convert_type<From,To> = To::convert_store_data(From::convert_load_data(x)): half/bf16 stores take a
 * FLOAT argument, so a double source is first narrowed to float (double rounding is intentional). */
static void store_elem(void* p, int dt, elem_t e, int src_dt)
{
  if (dt_is_float(dt)) {
    int src_is_double = (src_dt == DT_DOUBLE);
    switch (dt) {
      case DT_FLOAT: *(float*)p = src_is_double ? (float)e.d : e.f; break;
      case DT_DOUBLE: *(double*)p = src_is_double ? e.d : (double)e.f; break;
      case DT_HALF: *(uint16_t*)p = float_to_half(src_is_double ? (float)e.d : e.f); break;
      case DT_BF16: *(uint16_t*)p = float_to_bf16(src_is_double ? (float)e.d : e.f); break;
    }
  } else {
    switch (dt) {
      case DT_INT8: *(int8_t*)p = (int8_t)e.i; break;
      case DT_INT16: *(int16_t*)p = (int16_t)e.i; break;
      case DT_INT: *(int32_t*)p = (int32_t)e.i; break;
      case DT_INT64: *(int64_t*)p = e.i; break;
    }
  }
}

static int64_t load_index(const void* idx, int idx_dt, int64_t i)
{
  return idx_dt == DT_INT ? (int64_t)((const int32_t*)idx)[i] : ((const int64_t*)idx)[i];
}

/*
 * gather_func_kernel / gather_func_sub_warp_kernel, gather_scatter_func.cuh:289-314, :356-375:
 *   out[out_off + i*out_stride + c] = convert(table[tab_off + idx[i]*tab_stride + c]), idx[i] < 0 => skip.
 * `table` is the FLAT whole table (the oracle has no ranks: a gref only changes WHERE bytes live).
 * offsets / strides in elements.
 */
void oracle_gather(const void* table, int tab_dt, int64_t tab_stride, int64_t tab_off, int64_t cols,
                   const void* idx, int idx_dt, int64_t n,
                   void* out, int out_dt, int64_t out_stride, int64_t out_off)
{
  const int ts = dt_size(tab_dt), os = dt_size(out_dt);
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = load_index(idx, idx_dt, i);
    if (r < 0) continue;
    const char* src = (const char*)table + (tab_off + r * tab_stride) * ts;
    char* dst       = (char*)out + (out_off + i * out_stride) * os;
    if (tab_dt == out_dt) {
      memcpy(dst, src, (size_t)cols * ts);
    } else {
      for (int64_t c = 0; c < cols; ++c) store_elem(dst + c * os, out_dt, load_elem(src + c * ts, tab_dt), tab_dt);
    }
  }
}

/* scatter_func_kernel, gather_scatter_func.cuh:574-596: table[idx[i]] = convert(in[i]); idx<0 skipped;
 * duplicates: the reference races, the oracle applies them in order (last wins) -- tests only use
 * duplicates that carry identical data, as the reference's tests do. */
void oracle_scatter(const void* in, int in_dt, int64_t in_stride, int64_t in_off, int64_t cols,
                    const void* idx, int idx_dt, int64_t n,
                    void* table, int tab_dt, int64_t tab_stride, int64_t tab_off)
{
  const int is = dt_size(in_dt), ts = dt_size(tab_dt);
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = load_index(idx, idx_dt, i);
    if (r < 0) continue;
    const char* src = (const char*)in + (in_off + i * in_stride) * is;
    char* dst       = (char*)table + (tab_off + r * tab_stride) * ts;
    if (tab_dt == in_dt) {
      memcpy(dst, src, (size_t)cols * ts);
    } else {
      for (int64_t c = 0; c < cols; ++c) store_elem(dst + c * ts, tab_dt, load_elem(src + c * is, in_dt), in_dt);
    }
  }
}

/* element-wise conversion of a buffer (device_matrix_type_cast in the reference tests) */
void oracle_convert(const void* src, int src_dt, void* dst, int dst_dt, int64_t count)
{
  const int ss = dt_size(src_dt), ds = dt_size(dst_dt);
  for (int64_t i = 0; i < count; ++i)
    store_elem((char*)dst + i * ds, dst_dt, load_elem((const char*)src + i * ss, src_dt), src_dt);
}

/* The reference tests' closed-form table: every element of row r is convert(r & (2^(M+1)-1)), M =
 * mantissa bits of the dtype, ints take r truncated (embedding_test_utils.cu:197-238; the *97+1007
 * hash at :235 is overwritten at :236). */
void oracle_fill_test_pattern(void* table, int dt, int64_t first_row, int64_t rows, int64_t cols, int64_t stride)
{
  const int es = dt_size(dt);
  for (int64_t r = 0; r < rows; ++r) {
    int64_t g = first_row + r;
    for (int64_t c = 0; c < cols; ++c) {
      char* p = (char*)table + (r * stride + c) * es;
      switch (dt) {
        case DT_FLOAT: *(float*)p = (float)(g & ((1LL << 24) - 1)); break;
        case DT_DOUBLE: *(double*)p = (double)(g & ((1LL << 53) - 1)); break;
        case DT_HALF: *(uint16_t*)p = float_to_half((float)(g & ((1LL << 11) - 1))); break;
        case DT_BF16: *(uint16_t*)p = float_to_bf16((float)(g & ((1LL << 8) - 1))); break;
        case DT_INT8: *(int8_t*)p = (int8_t)g; break;
        case DT_INT16: *(int16_t*)p = (int16_t)g; break;
        case DT_INT: *(int32_t*)p = (int32_t)g; break;
        case DT_INT64: *(int64_t*)p = g; break;
      }
    }
  }
}

/* generate_rank_partition_strategy, memory_handle.cpp:1618-1635 + equal_partition_plan :2122-2128:
 * entries_per_rank = ceil(N / ws); rank r owns [min(r*e, N), min((r+1)*e, N)).  offsets has ws+1 slots. */
void oracle_partition(int64_t entries, int world_size, int64_t* offsets)
{
  int64_t per = (entries + world_size - 1) / world_size;
  for (int r = 0; r <= world_size; ++r) {
    int64_t o  = (int64_t)r * per;
    offsets[r] = o < entries ? o : entries;
  }
}

/* ---------------------------------------------------------------- sparse optimizers
 * One call = one optimizer step on the rows listed in `rows` (unique, as after dedup) with fp32
 * gradients g[k*gstride + c].  Arithmetic order copied from the kernels:
 *   SGD      embedding_optimizer_func.cu:212-223
 *   LazyAdam :386-418  (per-row beta1^t, beta2^t stored in b12[row*2 + {0,1}], updated BEFORE use)
 *   AdaGrad  :644-656
 *   RMSProp  :838-850
 * and they agree with CPUOptimizer in wholememory_embedding_gradient_apply_tests.cu:213-292. */
void oracle_sgd(float* w, int64_t wstride, int64_t dim, const int64_t* rows, int64_t n, const float* g, int64_t gstride,
                float weight_decay, float lr)
{
  for (int64_t k = 0; k < n; ++k) {
    float* wr       = w + rows[k] * wstride;
    const float* gr = g + k * gstride;
    for (int64_t c = 0; c < dim; ++c) {
      float gv = gr[c];
      float wv = wr[c];
      gv += weight_decay * wv;
      wv -= lr * gv;
      wr[c] = wv;
    }
  }
}

void oracle_lazy_adam(float* w, int64_t wstride, float* m, float* v, int64_t sstride, float* b12, int64_t dim,
                      const int64_t* rows, int64_t n, const float* g, int64_t gstride,
                      float weight_decay, float epsilon, float beta1, float beta2, int adam_w, float lr)
{
  for (int64_t k = 0; k < n; ++k) {
    int64_t r       = rows[k];
    float* wr       = w + r * wstride;
    float* mr       = m + r * sstride;
    float* vr       = v + r * sstride;
    const float* gr = g + k * gstride;
    float beta1t    = b12[r * 2 + 0] * beta1;
    float beta2t    = b12[r * 2 + 1] * beta2;
    for (int64_t c = 0; c < dim; ++c) {
      float gv = gr[c];
      float wv = wr[c];
      if (adam_w) {
        wv -= lr * weight_decay * wv;
      } else {
        gv = gv + weight_decay * wv;
      }
      float mm   = mr[c];
      float vv   = vr[c];
      mm         = beta1 * mm + (1 - beta1) * gv;
      vv         = beta2 * vv + (1 - beta2) * gv * gv;
      float mhat = mm / (1 - beta1t);
      float vhat = vv / (1 - beta2t);
      wv         = wv - lr * mhat / (sqrtf(vhat) + epsilon);
      mr[c]      = mm;
      vr[c]      = vv;
      wr[c]      = wv;
    }
    b12[r * 2 + 0] = beta1t;
    b12[r * 2 + 1] = beta2t;
  }
}

void oracle_adagrad(float* w, int64_t wstride, float* state_sum, int64_t sstride, int64_t dim,
                    const int64_t* rows, int64_t n, const float* g, int64_t gstride,
                    float weight_decay, float epsilon, float lr)
{
  for (int64_t k = 0; k < n; ++k) {
    float* wr       = w + rows[k] * wstride;
    float* sr       = state_sum + rows[k] * sstride;
    const float* gr = g + k * gstride;
    for (int64_t c = 0; c < dim; ++c) {
      float gv = gr[c];
      float wv = wr[c];
      gv       = gv + weight_decay * wv;
      float s  = sr[c];
      s        = s + gv * gv;
      wv       = wv - lr * gv / (sqrtf(s) + epsilon);
      sr[c]    = s;
      wr[c]    = wv;
    }
  }
}

void oracle_rmsprop(float* w, int64_t wstride, float* v, int64_t sstride, int64_t dim,
                    const int64_t* rows, int64_t n, const float* g, int64_t gstride,
                    float weight_decay, float epsilon, float alpha, float lr)
{
  for (int64_t k = 0; k < n; ++k) {
    float* wr       = w + rows[k] * wstride;
    float* vr       = v + rows[k] * sstride;
    const float* gr = g + k * gstride;
    for (int64_t c = 0; c < dim; ++c) {
      float gv = gr[c];
      float wv = wr[c];
      gv       = gv + weight_decay * wv;
      float vv = vr[c];
      vv       = alpha * vv + (1 - alpha) * gv * gv;
      wv       = wv - lr * gv / (sqrtf(vv) + epsilon);
      vr[c]    = vv;
      wr[c]    = wv;
    }
  }
}

/* dedup_indice_and_gradients, exchange_embeddings_nccl_func.cu:76-176: unique row ids ascending; the
 * gradients of duplicates are added one after another IN ARRIVAL ORDER (stable sort by id).
 * Returns the unique count; out_rows[u], out_g[u*dim + c]. */
typedef struct {
  int64_t id;
  int64_t pos;
} idpos_t;
static int idpos_cmp(const void* a, const void* b)
{
  const idpos_t *x = (const idpos_t*)a, *y = (const idpos_t*)b;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}
int64_t oracle_dedup_gradients(const int64_t* ids, int64_t n, const float* g, int64_t gstride, int64_t dim,
                               int64_t* out_rows, float* out_g)
{
  if (n == 0) return 0;
  idpos_t* a = (idpos_t*)malloc(sizeof(idpos_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    a[i].id  = ids[i];
    a[i].pos = i;
  }
  qsort(a, (size_t)n, sizeof(idpos_t), idpos_cmp);
  int64_t u = -1;
  for (int64_t i = 0; i < n; ++i) {
    const float* src = g + a[i].pos * gstride;
    if (i == 0 || a[i].id != a[i - 1].id) {
      ++u;
      out_rows[u] = a[i].id;
      for (int64_t c = 0; c < dim; ++c) out_g[u * dim + c] = src[c];
    } else {
      for (int64_t c = 0; c < dim; ++c) out_g[u * dim + c] += src[c];
    }
  }
  free(a);
  return u + 1;
}

/* ---------------------------------------------------------------- neighbor sampling
 * Random stream: PCG-XSH-RR 64/32 as RAFT's PCGenerator uses it (RAFT source is
 * not available here => stream parity UNPINNED).  init(seed, subsequence, offset=0):
 *   state = 0; inc = (subsequence << 1) | 1; step; state += seed; step.   next_u32 = xsh-rr output.
 *   next(int32) = next_u32 & 0x7fffffff  (hence "generate_random_positive_int", raft_random_gen.cu:27). */
typedef struct {
  uint64_t state, inc;
} pcg_t;
static uint32_t pcg_next_u32(pcg_t* g)
{
  uint64_t old   = g->state;
  g->state       = old * 6364136223846793005ULL + g->inc;
  uint32_t xs    = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot   = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((-rot) & 31u));
}
static void pcg_init(pcg_t* g, uint64_t seed, uint64_t subsequence)
{
  g->state = 0;
  g->inc   = (subsequence << 1u) | 1u;
  pcg_next_u32(g);
  g->state += seed;
  pcg_next_u32(g);
}

/* raft_random_gen.cu:27-71: output[i] = i-th positive int of (seed, subsequence) */
void oracle_random_positive_ints(uint64_t seed, uint64_t subsequence, int32_t* out, int64_t count)
{
  pcg_t g;
  pcg_init(&g, seed, subsequence);
  for (int64_t i = 0; i < count; ++i) out[i] = (int32_t)(pcg_next_u32(&g) & 0x7fffffffu);
}

/* random_sample_without_replacement_cpu_base, graph_sampling_test_utils.cu:306-321 ("pick and back-fill"):
 *   Q = [0..N);  for i in 0..M-1:  a[i] = Q[r[i]];  Q[r[i]] = Q[N-i-1];      r[i] in [0, N-i)
 * The GPU kernel (unweighted_sample_without_replacement_func.cuh:191-282: radix sort of (r,i) + pointer
 * jumping) is the parallel form of exactly this recurrence. */
void oracle_fisher_yates(const int32_t* r, int M, int N, int32_t* sample_pos)
{
  int32_t* q = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  for (int i = 0; i < N; ++i) q[i] = i;
  for (int i = 0; i < M; ++i) {
    sample_pos[i] = q[r[i]];
    q[r[i]]       = q[N - i - 1];
  }
  free(q);
}

/* (BLOCK_DIM, ITEMS_PER_THREAD) the reference picks from max_sample_count: func_array / warp_count_array,
 * unweighted_sample_without_replacement_func.cuh:423-458, indexed by (k-1)/32. */
static void sampler_shape(int k, int* block_dim, int* items)
{
  static const int wc[32] = {1, 1, 1, 2, 2, 2, 4, 4, 4, 4, 4, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
  static const int it[32] = {1, 2, 3, 2, 3, 3, 2, 2, 3, 3, 3, 3, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4};
  int f = (k - 1) / 32;
  if (f < 0) f = 0;
  if (f > 31) f = 31;
  *block_dim = wc[f] * 32;
  *items     = it[f];
}
void oracle_sampler_shape(int k, int* block_dim, int* items) { sampler_shape(k, block_dim, items); }

/*
 * wholegraph_csr_unweighted_sample_without_replacement, host driver
 * unweighted_sample_without_replacement_func.cuh:284-475 + kernel :126-282 (k <= 1024 path):
 *   offsets = exclusive scan of min(deg, k)
 *   deg <= k (or k <= 0: sample_all_kernel, sample_comm.cuh:24-58): copy all neighbours in CSR order
 *   else: thread t of the node's block owns generator subsequence (node_index*BLOCK_DIM + t) and draws
 *         ITEMS values; position i = j*BLOCK_DIM + t uses draw j of thread t; r[i] = draw % (deg - i)
 *         for i < k; selection = partial Fisher-Yates; outputs written in position order.
 * col ids / center ids are passed as int64 here (the oracle is dtype-agnostic for ids).
 * out_offsets has n+1 entries.  Returns the total sample count.
 */
int64_t oracle_unweighted_sample(const int64_t* row_ptr, const int64_t* col, const int64_t* centers, int64_t n, int k,
                                 uint64_t seed, int32_t* out_offsets, int64_t* out_dst, int32_t* out_center_local,
                                 int64_t* out_edge_gid)
{
  int64_t total = 0;
  for (int64_t c = 0; c < n; ++c) {
    int64_t deg    = row_ptr[centers[c] + 1] - row_ptr[centers[c]];
    out_offsets[c] = (int32_t)total;
    total += (k <= 0 || deg <= k) ? deg : k;
  }
  out_offsets[n] = (int32_t)total;
  if (out_dst == NULL) return total;
  int block_dim = 32, items = 1;
  if (k > 0) sampler_shape(k, &block_dim, &items);
  int32_t* r   = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
  int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
  for (int64_t c = 0; c < n; ++c) {
    int64_t start = row_ptr[centers[c]];
    int64_t deg   = row_ptr[centers[c] + 1] - start;
    int64_t o     = out_offsets[c];
    if (k <= 0 || deg <= k) {
      for (int64_t e = 0; e < deg; ++e) {
        out_dst[o + e] = col[start + e];
        if (out_center_local) out_center_local[o + e] = (int32_t)c;
        if (out_edge_gid) out_edge_gid[o + e] = start + e;
      }
      continue;
    }
    for (int t = 0; t < block_dim; ++t) {
      pcg_t g;
      pcg_init(&g, seed, (uint64_t)(c * block_dim + t));
      for (int j = 0; j < items; ++j) {
        int32_t draw = (int32_t)(pcg_next_u32(&g) & 0x7fffffffu);
        int i        = j * block_dim + t;
        if (i < k) r[i] = (int32_t)(draw % (int32_t)(deg - i));
      }
    }
    oracle_fisher_yates(r, k, (int)deg, pos);
    for (int i = 0; i < k; ++i) {
      out_dst[o + i] = col[start + pos[i]];
      if (out_center_local) out_center_local[o + i] = (int32_t)c;
      if (out_edge_gid) out_edge_gid[o + i] = start + pos[i];
    }
  }
  free(r);
  free(pos);
  return total;
}

/* ---------------------------------------------------------------- timed CPU baseline
 * Multithreaded same-dtype row gather: one contiguous slice of the index batch per thread, memcpy per
 * row (BASELINE.md section 3 "CPU" line).  Result identical to oracle_gather. */
typedef struct {
  const char* table;
  const int64_t* idx;
  char* out;
  int64_t row_bytes, tab_stride_bytes, out_stride_bytes, begin, end;
} mt_job_t;
static void* mt_worker(void* arg)
{
  mt_job_t* j = (mt_job_t*)arg;
  for (int64_t i = j->begin; i < j->end; ++i) {
    int64_t r = j->idx[i];
    if (r < 0) continue;
    memcpy(j->out + i * j->out_stride_bytes, j->table + r * j->tab_stride_bytes, (size_t)j->row_bytes);
  }
  return NULL;
}
void oracle_gather_mt(const void* table, int64_t tab_stride_bytes, int64_t row_bytes, const int64_t* idx, int64_t n,
                      void* out, int64_t out_stride_bytes, int threads)
{
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  mt_job_t job[256];
  for (int t = 0; t < threads; ++t) {
    job[t].table            = (const char*)table;
    job[t].idx              = idx;
    job[t].out              = (char*)out;
    job[t].row_bytes        = row_bytes;
    job[t].tab_stride_bytes = tab_stride_bytes;
    job[t].out_stride_bytes = out_stride_bytes;
    job[t].begin            = n * t / threads;
    job[t].end              = n * (t + 1) / threads;
    pthread_create(&tid[t], NULL, mt_worker, &job[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
}
