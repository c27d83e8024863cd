/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see wm_oracle.c).  Weighted neighbor sampling restated in plain C.
 *
 * Follows reference cpp/src/wholegraph_ops/weighted_sample_without_replacement_func.cuh:45-63 (gen_key_from_weight),
 * :219-297 (A-Res: keep the k largest keys; thread t of a B-thread CTA owns generator subsequence center*B + t and the
 * neighbours t, t+B, ...; B = 128, or 256 when k > RAFT kMaxCapacity = 256) and the reference's own CPU model
 * cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:552-668.  Keys are computed in float like the device code (the
 * reference's CPU model uses double and relies on comparing sorted outputs).  Random stream: RAFT PCGenerator restated,
 * parity UNPINNED (wm_oracle.c header).  Ties: larger key first, then smaller neighbour index.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
  uint64_t state, inc;
} wpcg_t;
static uint32_t wpcg_next(wpcg_t* g)
{
  uint64_t old = g->state;
  g->state     = old * 6364136223846793005ULL + g->inc;
  uint32_t xs  = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((-rot) & 31u));
}
static void wpcg_init(wpcg_t* g, uint64_t seed, uint64_t subsequence)
{
  g->state = 0;
  g->inc   = (subsequence << 1u) | 1u;
  wpcg_next(g);
  g->state += seed;
  wpcg_next(g);
}

static float key_from_weight(double weight, wpcg_t* g)
{
  float u = (float)(wpcg_next(g) >> 8) / 16777216.0f;
  u       = -(0.5f + 0.5f * u);
  uint64_t r2 = 0;
  int rounds  = -1;
  do {
    uint64_t lo = wpcg_next(g);
    uint64_t hi = wpcg_next(g);
    r2          = lo | (hi << 32);
    ++rounds;
  } while (r2 == 0);
  int one_bit = __builtin_clzll(r2) + rounds * 64;
  u *= exp2f(-(float)one_bit);
  return (log1pf(u) / logf(2.0f)) * (1.0f / (float)weight);
}

/* The reference's HOST replay of the key stream, raft_random_gen.cu:73-119: same draws as key_from_weight with weight 1,
 * but the final step is DOUBLE arithmetic -- log1p(u) / log(2.0), rounded once to float -- not the device kernels'
 * log1pf(u) / logf(2.0).  Pinned against the reference source compiled for the CPU (tests/test_ref_host_random.py). */
void oracle_exponential_negative_floats(uint64_t seed, uint64_t subsequence, float* out, int64_t count)
{
  wpcg_t g;
  wpcg_init(&g, seed, subsequence);
  for (int64_t i = 0; i < count; ++i) {
    float u = (float)(wpcg_next(&g) >> 8) / 16777216.0f;
    u       = (float)-(0.5 + 0.5 * (double)u);
    uint64_t r2 = 0;
    int rounds  = -1;
    do {
      uint64_t lo = wpcg_next(&g);
      uint64_t hi = wpcg_next(&g);
      r2          = lo | (hi << 32);
      ++rounds;
    } while (r2 == 0);
    int one_bit = __builtin_clzll(r2) + rounds * 64;
    u           = (float)((double)u * pow(2.0, (double)-one_bit));
    out[i]      = (float)(log1p((double)u) / log(2.0));
  }
}

typedef struct {
  float key;
  int idx;
} cand_t;
static int cand_cmp(const void* a, const void* b)
{
  const cand_t *x = (const cand_t*)a, *y = (const cand_t*)b;
  if (x->key != y->key) return x->key > y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* weights are passed as double; weight_is_float says whether the table stored float (the cast chain is then
 * float -> float, else double -> float, exactly like `(float)weight` on the device).  keys_out (optional) receives, per
 * center, the k-th and (k+1)-th best keys so a test can tell a real mismatch from a last-ulp tie. */
int64_t oracle_weighted_sample(const int64_t* row_ptr, const int64_t* col, const double* weights, int weight_is_float,
                               const int64_t* centers, int64_t n, int k, uint64_t seed, int32_t* out_offsets, int64_t* out_dst,
                               int32_t* out_center_local, int64_t* out_edge_gid, float* margin_out)
{
  int64_t total = 0;
  for (int64_t c = 0; c < n; ++c) {
    int64_t deg    = row_ptr[centers[c] + 1] - row_ptr[centers[c]];
    out_offsets[c] = (int32_t)total;
    total += (k <= 0 || deg <= k) ? deg : k;
  }
  out_offsets[n] = (int32_t)total;
  if (out_dst == NULL) return total;
  const int B = k > 256 ? 256 : 128;
  for (int64_t c = 0; c < n; ++c) {
    int64_t start = row_ptr[centers[c]];
    int64_t deg   = row_ptr[centers[c] + 1] - start;
    int64_t o     = out_offsets[c];
    if (margin_out) margin_out[c] = INFINITY;
    if (k <= 0 || deg <= k) {
      for (int64_t e = 0; e < deg; ++e) {
        out_dst[o + e] = col[start + e];
        if (out_center_local) out_center_local[o + e] = (int32_t)c;
        if (out_edge_gid) out_edge_gid[o + e] = start + e;
      }
      continue;
    }
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)deg);
    for (int t = 0; t < B && t < deg; ++t) {
      wpcg_t g;
      wpcg_init(&g, seed, (uint64_t)(c * B + t));
      for (int64_t idx = t; idx < deg; idx += B) {
        double w = weights[start + idx];
        if (weight_is_float) w = (double)(float)w;
        cand[idx].key = key_from_weight(w, &g);
        cand[idx].idx = (int)idx;
      }
    }
    qsort(cand, (size_t)deg, sizeof(cand_t), cand_cmp);
    for (int i = 0; i < k; ++i) {
      out_dst[o + i] = col[start + cand[i].idx];
      if (out_center_local) out_center_local[o + i] = (int32_t)c;
      if (out_edge_gid) out_edge_gid[o + i] = start + cand[i].idx;
    }
    if (margin_out) margin_out[c] = fabsf(cand[k - 1].key - cand[k].key) / fmaxf(fabsf(cand[k].key), 1e-30f);
    free(cand);
  }
  return total;
}
