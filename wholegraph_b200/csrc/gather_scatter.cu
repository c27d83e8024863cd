/*
 * Host-side planning + launch for the row gather/scatter kernels (gather_scatter.cuh).
 * Replaces the reference's gather_temp_func / scatter_temp_func launch logic
 * (cpp/src/wholememory_ops/functions/gather_scatter_func.cuh:378-517, :600-661) and its
 * dtype-pair registry (register.hpp): same accepted dtype pairs plus bf16.
 */
#ifdef WG_DEV_KNOBS
#include "gather_bulk.cuh" /* copy-engine variant: measured slower everywhere on B200, kept for sweeps */
#endif
#include "gather_scatter.cuh"
#include "ops_internal.hpp"

#include <algorithm>

namespace wm {

namespace {

/* Launch shape, measured on B200 with tools/rowmove_lab.cu (profiles/README.md, round 2; 1,048,576 rows per launch):
 *   - 128-thread CTAs, ONE batch per warp, non-persistent grid (the hardware CTA scheduler walks the index array in
 *     order); a warp batch of ~4 KiB (R = 4096 / row_bytes rows): 256 B rows 0.919 -> 0.960 of HBM, 512 B 0.934 -> 0.987,
 *     1 KiB 0.979 -> 0.998.  Smaller batches (2 KiB) lose 3-9 %, a persistent grid-stride loop 5-9 %.
 *   - 256-bit accesses (LDG.E.256 / STG.E.256) wherever every address, stride and the row size are multiples of 32 B:
 *     +2-4 % (1 KiB rows: 1.017 of the measured copy bandwidth).
 *   - programmatic dependent launch: back-to-back calls overlap their launch ramp with the predecessor's tail, +1.5-2.5 %
 *     on the short kernels (256 B rows: 0.952 -> 0.976).
 * The copy-engine (cp.async.bulk) kernel measured 0.84-0.89 on the same shapes and is built only with -DWG_DEV_KNOBS. */
constexpr int kThreads     = 128;
constexpr int kUnroll      = 4;
constexpr int kBatchBytes  = 4096;

inline int pow2_divisor(uint64_t v, int cap)
{
  int a = cap;
  while (a > 1 && (v % (uint64_t)a) != 0) a >>= 1;
  return a;
}

/* Sweep knobs for the GPU box (WG_UNROLL, WG_THREADS, WG_BATCH_ROWS, WG_BLOCKS_PER_SM, WG_CACHE_POLICY, WG_GRID_MODE,
 * WG_VEC, WG_PDL, WG_BULK*): compiled in only with -DWG_DEV_KNOBS (make DEV_KNOBS=1); the shipped library ignores them. */
int env_int(const char* name, int dflt)
{
#ifdef WG_DEV_KNOBS
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
#ifdef WG_DEV_KNOBS
int tuned_unroll()
{
  static const int u = env_int("WG_UNROLL", kUnroll);
  return u;
}
#endif

int tuned_threads()
{
  static const int t = env_int("WG_THREADS", kThreads);
  return t;
}

/* One launch path for every instantiation: programmatic stream serialization when allowed (the kernels call
 * griddepcontrol.wait before their first dependent read), plain launch otherwise or when the attribute is refused. */
template <typename K, typename... Args>
void launch_row_kernel(K kernel, int grid, int threads, cudaStream_t s, Args... args)
{
  static const int pdl_mode = env_int("WG_PDL", 1);
  static bool pdl_ok        = pdl_mode != 0;
  if (pdl_ok) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim  = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.stream   = s;
    cudaLaunchAttribute attr;
    attr.id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                       = &attr;
    cfg.numAttrs                                    = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e == cudaSuccess) return;
    if (e != cudaErrorNotSupported && e != cudaErrorInvalidValue) WM_CUDA(e);
    (void)cudaGetLastError(); /* this driver / stream does not take the attribute: never try again */
    pdl_ok = false;
  }
  kernel<<<grid, threads, 0, s>>>(args...);
}

template <typename IdxT, int VEC, bool GATHER>
void launch_vec(const table_ref& t, const row_geom& g, const void* idx, int64_t n, char* dense, int grid, cudaStream_t s)
{
  const IdxT* ip = static_cast<const IdxT*>(idx);
#ifdef WG_DEV_KNOBS
  if constexpr (VEC == 16) {
    switch (tuned_unroll()) {
      case 2: launch_row_kernel(row_move_vec_kernel<IdxT, VEC, GATHER, 2>, grid, tuned_threads(), s, t, g, ip, n, dense); return;
      case 8: launch_row_kernel(row_move_vec_kernel<IdxT, VEC, GATHER, 8>, grid, tuned_threads(), s, t, g, ip, n, dense); return;
      default: break;
    }
  }
#endif
  launch_row_kernel(row_move_vec_kernel<IdxT, VEC, GATHER, kUnroll>, grid, tuned_threads(), s, t, g, ip, n, dense);
}

template <typename IdxT, bool GATHER>
void launch_vec_w(int vec, const table_ref& t, const row_geom& g, const void* idx, int64_t n, char* dense, int grid, cudaStream_t s)
{
  switch (vec) {
    case 32: launch_vec<IdxT, 32, GATHER>(t, g, idx, n, dense, grid, s); break;
    case 16: launch_vec<IdxT, 16, GATHER>(t, g, idx, n, dense, grid, s); break;
    case 8: launch_vec<IdxT, 8, GATHER>(t, g, idx, n, dense, grid, s); break;
    case 4: launch_vec<IdxT, 4, GATHER>(t, g, idx, n, dense, grid, s); break;
    case 2: launch_vec<IdxT, 2, GATHER>(t, g, idx, n, dense, grid, s); break;
    default: launch_vec<IdxT, 1, GATHER>(t, g, idx, n, dense, grid, s); break;
  }
}

int vec_blocks_per_sm()
{
  static int occ = [] {
    int o         = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, row_move_vec_kernel<int64_t, 32, true, kUnroll>, tuned_threads(), 0);
    if (e != cudaSuccess || o <= 0) {
      (void)cudaGetLastError();
      o = 8;
    }
    o = env_int("WG_BLOCKS_PER_SM", o);
    return o;
  }();
  return occ;
}

/* magic multiplier for floor(w / d) = (w * magic) >> 40, exact while w < 2^20 and d < 2^20.  Only batches of more than one
 * row use it, and plan() batches rows only up to kBatchBytes (<= 4096 units per batch); longer rows travel one per warp
 * and are limited only by the 32-bit unit counter (2^31 units: 2 GiB rows at 1-byte units, 64 GiB at 32-byte units). */
void set_units(row_geom* g, int64_t units_per_row)
{
  WM_EXPECT(units_per_row > 0 && units_per_row < ((int64_t)1 << 31), WHOLEMEMORY_NOT_SUPPORTED,
            "row too long for the gather/scatter kernels (%ld units)", (long)units_per_row);
  g->units_per_row = (int)units_per_row;
  g->div_magic     = (((uint64_t)1 << 40) + (uint64_t)units_per_row - 1) / (uint64_t)units_per_row;
}

/* rows per warp batch + grid size */
void plan(int64_t n, int64_t row_bytes, int sms, int blocks_per_sm, int* batch_rows, int* grid)
{
  int total_sms = sm_count();
  if (sms <= 0 || sms > total_sms) sms = total_sms;
  const int wpc       = tuned_threads() / 32;
  int64_t max_grid    = (int64_t)sms * blocks_per_sm;
  int64_t total_warps = max_grid * wpc;
  int R               = 32;
  while (R > 1 && (int64_t)R * row_bytes > kBatchBytes) R >>= 1;
  while (R > 1 && n / R < total_warps * 2) R >>= 1; /* small calls: spread the rows over the whole GPU */
  static const int forced_rows = env_int("WG_BATCH_ROWS", 0);
  if (forced_rows > 0 && (int64_t)forced_rows * row_bytes <= 16384) R = forced_rows; /* keeps R * units inside the magic's range */
  int64_t nbatch = (n + R - 1) / R;
  int64_t need   = (nbatch + wpc - 1) / wpc;
  *batch_rows    = R;
  *grid          = (int)std::max<int64_t>(1, std::min(max_grid, need));
  /* Unless the caller restricts the SM budget (gather_sms), launch ONE BATCH PER WARP and let the hardware CTA scheduler
   * walk the index array in order; a budget turns the same kernel into a persistent grid of sms * blocks_per_sm CTAs. */
  static const int grid_mode = env_int("WG_GRID_MODE", 1);
  if (grid_mode == 1 && sms == total_sms) *grid = (int)std::min<int64_t>(need, 0x7fffffff);
}

#ifdef WG_DEV_KNOBS
/* TMA-bulk variant (gather_bulk.cuh).  Returns false when the shape does not qualify. */
template <typename IdxT, bool GATHER>
bool launch_bulk(const table_ref& t, row_geom g, const void* idx, int64_t n, char* dense, int64_t row_bytes, int sms, cudaStream_t s)
{
  static const int slot_kb = env_int("WG_BULK_SLOT_KB", 4); /* 4 KiB slots: 754 vs 744 GB/s per GPU at 8 GPUs, 0.377 vs 0.390 ms local */
  int R = 32;
  while (R > 1 && (int64_t)R * row_bytes > (int64_t)slot_kb * 1024) R >>= 1;
  size_t smem = 128 + (size_t)kBulkWarps * kBulkStages * R * row_bytes;
  if (smem > 220 * 1024) return false;
  auto kernel = row_move_bulk_kernel<IdxT, GATHER>;
  static bool attr_set = false; /* per instantiation */
  if (!attr_set) {
    WM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBulkWarps * 32, smem) != cudaSuccess || occ < 1) {
    (void)cudaGetLastError();
    occ = 1;
  }
  int total_sms = sm_count();
  if (sms <= 0 || sms > total_sms) sms = total_sms;
  g.batch_rows   = R;
  int64_t nbatch = (n + R - 1) / R;
  int64_t need   = (nbatch + kBulkWarps - 1) / kBulkWarps;
  int grid       = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * occ, need));
  kernel<<<grid, kBulkWarps * 32, smem, s>>>(t, g, static_cast<const IdxT*>(idx), n, dense, (int)row_bytes);
  return true;
}
#endif

}  // namespace

table_ref make_flat_table_ref(void* base)
{
  table_ref t{};
  t.mode    = table_ref::FLAT;
  t.nranks  = 1;
  t.base[0] = static_cast<char*>(base);
  return t;
}

void row_move(bool gather,
              const table_ref& tref,
              const wholememory_matrix_description_t& td,
              const void* indices,
              const wholememory_array_description_t& idx_desc,
              void* dense,
              const wholememory_matrix_description_t& dd,
              cudaStream_t stream,
              int sms)
{
  require_cuda(gather ? "wholememory_gather" : "wholememory_scatter");
  const bool t_float = wholememory_dtype_is_floating_number(td.dtype);
  const bool d_float = wholememory_dtype_is_floating_number(dd.dtype);
  WM_EXPECT(t_float || wholememory_dtype_is_integer_number(td.dtype), WHOLEMEMORY_LOGIC_ERROR, "table dtype %d unsupported", (int)td.dtype);
  WM_EXPECT(d_float || wholememory_dtype_is_integer_number(dd.dtype), WHOLEMEMORY_LOGIC_ERROR, "dense dtype %d unsupported", (int)dd.dtype);
  /* reference gather_func.cu:78-81 */
  WM_EXPECT(t_float == d_float, WHOLEMEMORY_LOGIC_ERROR,
            "table and %s must both be floating point or both be integer", gather ? "output" : "input");
  WM_EXPECT(idx_desc.dtype == WHOLEMEMORY_DT_INT || idx_desc.dtype == WHOLEMEMORY_DT_INT64, WHOLEMEMORY_LOGIC_ERROR,
            "indices must be int32 or int64");
  const int64_t n = idx_desc.size;
  WM_EXPECT(dd.sizes[0] == n, WHOLEMEMORY_LOGIC_ERROR, "%s rows=%ld but indice_count=%ld",
            gather ? "output" : "input", (long)dd.sizes[0], (long)n);
  WM_EXPECT(dd.sizes[1] == td.sizes[1], WHOLEMEMORY_LOGIC_ERROR, "row width mismatch: table %ld vs %ld",
            (long)td.sizes[1], (long)dd.sizes[1]);
  if (n == 0 || td.sizes[1] == 0) return;
  WM_EXPECT(indices != nullptr && dense != nullptr, WHOLEMEMORY_INVALID_INPUT, "null indices / data pointer");

  const int64_t et = (int64_t)wholememory_dtype_get_element_size(td.dtype);
  const int64_t ed = (int64_t)wholememory_dtype_get_element_size(dd.dtype);
  const bool idx64 = idx_desc.dtype == WHOLEMEMORY_DT_INT64;
  const char* idx_ptr = static_cast<const char*>(indices) + idx_desc.storage_offset * (idx64 ? 8 : 4);
  char* dense_ptr     = static_cast<char*>(dense) + dd.storage_offset * ed;

  row_geom g{};
  g.table_offset_bytes = td.storage_offset * et;
  g.table_stride_bytes = td.stride * et;
  g.dense_stride_bytes = dd.stride * ed;
  /* LDG kernel: local rows stream past L1 (no_allocate); possibly-remote rows use plain L1-allocating loads,
   * which measured 4.5 % faster over NVLink (657 vs 629 GB/s). WG_CACHE_POLICY overrides. */
  static const int cache_policy = env_int("WG_CACHE_POLICY", -1);
  g.policy                      = cache_policy >= 0 ? cache_policy : (tref.has_remote ? 2 : 0);

  /* alignment shared by both sides, in bytes of each side's element */
  uint64_t t_bits = (uint64_t)g.table_offset_bytes | (uint64_t)g.table_stride_bytes;
  if (tref.mode == table_ref::FLAT) t_bits |= reinterpret_cast<uint64_t>(tref.base[0]);
  uint64_t d_bits = reinterpret_cast<uint64_t>(dense_ptr) | (uint64_t)g.dense_stride_bytes;

  if (td.dtype == dd.dtype) {
    const int64_t row_bytes = td.sizes[1] * et;
    /* widest unit: 32 bytes (sm_100's 256-bit global accesses) where every address, stride and the row size allow it,
     * else 16, 8, ... 1 */
    static const int max_vec = env_int("WG_VEC", 32);
    int vec                  = pow2_divisor(t_bits | d_bits | (uint64_t)row_bytes, max_vec);
    g.row_elems              = (int)td.sizes[1];
    set_units(&g, row_bytes / vec);
    int grid                 = 1;
#ifdef WG_DEV_KNOBS
    static const int bulk_mode = env_int("WG_BULK", 0);
    if (bulk_mode != 0 && vec >= 16 && row_bytes >= 64) {
      set_units(&g, row_bytes / 16);
      bool done = gather ? (idx64 ? launch_bulk<int64_t, true>(tref, g, idx_ptr, n, dense_ptr, row_bytes, sms, stream)
                                  : launch_bulk<int32_t, true>(tref, g, idx_ptr, n, dense_ptr, row_bytes, sms, stream))
                         : (idx64 ? launch_bulk<int64_t, false>(tref, g, idx_ptr, n, dense_ptr, row_bytes, sms, stream)
                                  : launch_bulk<int32_t, false>(tref, g, idx_ptr, n, dense_ptr, row_bytes, sms, stream));
      if (done) {
        WM_CUDA(cudaGetLastError());
        return;
      }
      set_units(&g, row_bytes / vec);
    }
#endif
    plan(n, row_bytes, sms, vec_blocks_per_sm(), &g.batch_rows, &grid);
    if (gather) {
      if (idx64) launch_vec_w<int64_t, true>(vec, tref, g, idx_ptr, n, dense_ptr, grid, stream);
      else launch_vec_w<int32_t, true>(vec, tref, g, idx_ptr, n, dense_ptr, grid, stream);
    } else {
      if (idx64) launch_vec_w<int64_t, false>(vec, tref, g, idx_ptr, n, dense_ptr, grid, stream);
      else launch_vec_w<int32_t, false>(vec, tref, g, idx_ptr, n, dense_ptr, grid, stream);
    }
  } else {
    /* elements per lane: both vectors <= 16 bytes and aligned (reference :215-251, :395-397) */
    int cap   = (int)(16 / std::max(et, ed));
    int a_t   = pow2_divisor(t_bits / (uint64_t)et | (uint64_t)td.sizes[1], cap);
    int a_d   = pow2_divisor(d_bits / (uint64_t)ed | (uint64_t)td.sizes[1], cap);
    /* t_bits/d_bits are multiples of the element size by construction of the descriptors */
    int align = std::min(a_t, a_d);
    g.row_elems = (int)td.sizes[1];
    set_units(&g, td.sizes[1] / align);
    int grid    = 1;
    plan(n, td.sizes[1] * std::max(et, ed), sms, cvt_blocks_per_sm(), &g.batch_rows, &grid);
    cvt_launch_fn fn = t_float ? find_float_cvt(td.dtype, dd.dtype) : find_int_cvt(td.dtype, dd.dtype);
    WM_EXPECT(fn != nullptr, WHOLEMEMORY_LOGIC_ERROR, "no conversion kernel for dtype %d -> %d", (int)td.dtype, (int)dd.dtype);
    fn(gather, tref, g, idx_ptr, idx64, n, dense_ptr, align, grid, stream);
  }
  WM_CUDA(cudaGetLastError());
  static const bool debug_sync = getenv("WM_DEBUG_SYNC") != nullptr; /* reference cuda_macros.cpp:27-66 */
  if (debug_sync) WM_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace wm
