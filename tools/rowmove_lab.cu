/*
 * rowmove_lab: one-process laboratory for the row gather / scatter kernels of wholegraph_b200/csrc/gather_scatter.cuh
 * (the SAME kernel source the library ships, included directly), built for kernel experiments that would otherwise cost
 * one process per configuration:
 *
 *   - launch-shape sweeps (threads per CTA, rows per warp batch, loads in flight, 128/256-bit accesses, cache policy,
 *     persistent vs one-batch-per-warp grid, programmatic dependent launch, compile-time row width, copy-engine kernel)
 *   - access-pattern diagnostics (uniform random vs sequential indices)
 *   - NVLink experiments inside ONE process (so ncu can wrap it): device 0 gathers rows that live on device 1
 *     ("uni"), or both devices gather from each other at the same time ("bidir"), reads or peer stores, with NVML's
 *     NVLink data/raw byte counters read around every timed loop.
 *
 * Build:  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Iinclude -Iwholegraph_b200/csrc \
 *              tools/rowmove_lab.cu -o wholegraph_b200/lib/rowmove_lab -ldl
 * Run  :  rowmove_lab --row-bytes 256 --rows 20000000 --n 1048576 --mode local --set small
 * Output: one line per variant: name, ms per launch, GB/s out, algorithmic GB/s (2*row+8 per row), fraction of --peak.
 * Timing: CUDA events on the launching stream around `--iters` back-to-back launches after `--warmup`, 8 index batches
 * cycled, table larger than L2.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvml.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "gather_bulk.cuh"
#include "gather_scatter.cuh"

#define CK(x)                                                                                    \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));         \
      exit(2);                                                                                   \
    }                                                                                            \
  } while (0)

using namespace wm;

struct variant {
  std::string name;
  int threads = 256, R = 0 /* 0 = library rule */, unroll = 4, vec = 16, policy = -1 /* -1 = library rule */;
  int persistent = 0, bulk = 0, pdl = 0, vn_ct = 0, slot_kb = 4;
};

struct side { /* everything one device needs */
  int dev = 0;
  char* table = nullptr;     /* this device's shard */
  char* out = nullptr;       /* dense side */
  int64_t* idx[8] = {};
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  table_ref tref{};
};

static int g_sms = 148;

template <typename K>
static void launch_k(K kernel, int grid, int threads, size_t smem, cudaStream_t s, bool pdl, table_ref t, row_geom g, const int64_t* idx,
                     int64_t n, char* dense)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim          = dim3((unsigned)grid);
  cfg.blockDim         = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = s;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                        = at;
  cfg.numAttrs                                     = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, kernel, t, g, idx, n, dense));
}

template <int VEC, bool GATHER>
static void launch_variant(const variant& v, const table_ref& t, row_geom g, const int64_t* idx, int64_t n, char* dense, int64_t row_bytes,
                           cudaStream_t s)
{
  int R = v.R;
  if (R == 0) {
    R = 32;
    while (R > 1 && (int64_t)R * row_bytes > 16384) R >>= 1;
  }
  g.batch_rows    = R;
  int64_t nbatch  = (n + R - 1) / R;
  const int wpc   = v.threads / 32;
  int64_t need    = (nbatch + wpc - 1) / wpc;
#define OCC(K)                                                                                 \
  [&] {                                                                                        \
    int o = 0;                                                                                 \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, K, v.threads, 0));                    \
    return o;                                                                                  \
  }()
#define GO(U, VN)                                                                                                          \
  do {                                                                                                                     \
    auto k   = row_move_vec_kernel<int64_t, VEC, GATHER, U, VN>;                                                            \
    int grid = v.persistent ? (int)std::min<int64_t>(need, (int64_t)g_sms * OCC(k)) : (int)need;                            \
    launch_k(k, grid, v.threads, 0, s, v.pdl != 0, t, g, idx, n, dense);                                                    \
  } while (0)
  if (v.vn_ct == 16 && VEC == 16) {
    if (v.unroll == 8) GO(8, 16); else if (v.unroll == 2) GO(2, 16); else GO(4, 16);
  } else if (v.vn_ct == 32 && VEC == 16) {
    if (v.unroll == 8) GO(8, 32); else GO(4, 32);
  } else if (v.vn_ct == 64 && VEC == 16) {
    if (v.unroll == 8) GO(8, 64); else GO(4, 64);
  } else {
    if (v.unroll == 8) GO(8, 0); else if (v.unroll == 2) GO(2, 0); else if (v.unroll == 1) GO(1, 0); else GO(4, 0);
  }
#undef GO
#undef OCC
}

template <bool GATHER>
static void launch_bulk_variant(const variant& v, const table_ref& t, row_geom g, const int64_t* idx, int64_t n, char* dense, int64_t row_bytes,
                                cudaStream_t s)
{
  int R = 32;
  while (R > 1 && (int64_t)R * row_bytes > (int64_t)v.slot_kb * 1024) R >>= 1;
  size_t smem = 128 + (size_t)kBulkWarps * kBulkStages * R * row_bytes;
  auto k      = row_move_bulk_kernel<int64_t, GATHER>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  int occ = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kBulkWarps * 32, smem));
  g.batch_rows   = R;
  int64_t nbatch = (n + R - 1) / R;
  int64_t need   = (nbatch + kBulkWarps - 1) / kBulkWarps;
  int grid       = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)g_sms * occ, need));
  k<<<grid, kBulkWarps * 32, smem, s>>>(t, g, idx, n, dense, (int)row_bytes);
}

static void launch(const variant& v, bool gather, const side& sd, int batch, int64_t n, int64_t row_bytes, bool remote)
{
  row_geom g{};
  g.table_stride_bytes = row_bytes;
  g.dense_stride_bytes = row_bytes;
  g.row_elems          = (int)(row_bytes / 4);
  g.policy             = v.policy >= 0 ? v.policy : (remote ? 2 : 0);
  int vec              = v.vec;
  g.units_per_row      = (int)(row_bytes / vec);
  g.div_magic          = (((uint64_t)1 << 40) + (uint64_t)g.units_per_row - 1) / (uint64_t)g.units_per_row;
  if (v.bulk) {
    if (gather) launch_bulk_variant<true>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
    else launch_bulk_variant<false>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
  } else if (vec == 32) {
    if (gather) launch_variant<32, true>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
    else launch_variant<32, false>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
  } else {
    if (gather) launch_variant<16, true>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
    else launch_variant<16, false>(v, sd.tref, g, sd.idx[batch], n, sd.out, row_bytes, sd.s);
  }
  CK(cudaGetLastError());
}

/* ---- NVML NVLink counters (resolved at run time; absent => zeros) ---- */
struct nvl_counters {
  unsigned long long data_tx = 0, data_rx = 0, raw_tx = 0, raw_rx = 0; /* KiB, summed over links */
};
static void* g_nvml          = nullptr;
static nvmlDevice_t g_nvdev[2];
static decltype(&nvmlDeviceGetFieldValues) p_fields = nullptr;
static bool nvml_open(int ndev)
{
  g_nvml = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!g_nvml) return false;
  auto init   = (nvmlReturn_t(*)())dlsym(g_nvml, "nvmlInit_v2");
  auto byidx  = (nvmlReturn_t(*)(unsigned, nvmlDevice_t*))dlsym(g_nvml, "nvmlDeviceGetHandleByPciBusId_v2");
  auto bypci  = (nvmlReturn_t(*)(const char*, nvmlDevice_t*))dlsym(g_nvml, "nvmlDeviceGetHandleByPciBusId_v2");
  p_fields    = (decltype(p_fields))dlsym(g_nvml, "nvmlDeviceGetFieldValues");
  (void)byidx;
  if (!init || !bypci || !p_fields || init() != NVML_SUCCESS) return false;
  for (int d = 0; d < ndev; ++d) {
    char bus[64];
    CK(cudaDeviceGetPCIBusId(bus, sizeof(bus), d));
    if (bypci(bus, &g_nvdev[d]) != NVML_SUCCESS) return false;
  }
  return true;
}
static nvl_counters nvl_read(int d)
{
  nvl_counters c;
  if (!p_fields) return c;
  nvmlFieldValue_t f[4];
  memset(f, 0, sizeof(f));
  f[0].fieldId = NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX;
  f[1].fieldId = NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX;
  f[2].fieldId = NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX;
  f[3].fieldId = NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX;
  for (auto& x : f) x.scopeId = UINT_MAX; /* sum over all links */
  if (p_fields(g_nvdev[d], 4, f) != NVML_SUCCESS) return c;
  unsigned long long* out[4] = {&c.data_tx, &c.data_rx, &c.raw_tx, &c.raw_rx};
  for (int i = 0; i < 4; ++i)
    if (f[i].nvmlReturn == NVML_SUCCESS) *out[i] = f[i].value.ullVal;
  return c;
}

static std::vector<variant> variant_set(const std::string& set, int64_t row_bytes)
{
  std::vector<variant> vs;
  auto add = [&](const char* name, auto fn) {
    variant v;
    v.name = name;
    fn(v);
    vs.push_back(v);
  };
  add("default", [](variant&) {});
  if (set == "default") return vs;
  add("pdl", [](variant& v) { v.pdl = 1; });
  if (set == "small" || set == "all") {
    int vn = (int)(row_bytes / 16);
    bool ct = vn == 16 || vn == 32 || vn == 64;
    for (int th : {64, 128, 512}) add(("threads" + std::to_string(th)).c_str(), [&](variant& v) { v.threads = th; });
    for (int R : {4, 8, 16, 32}) add(("R" + std::to_string(R)).c_str(), [&](variant& v) { v.R = R; });
    for (int u : {1, 2, 8}) add(("unroll" + std::to_string(u)).c_str(), [&](variant& v) { v.unroll = u; });
    add("unroll8_R32_th128", [](variant& v) { v.unroll = 8; v.R = 32; v.threads = 128; });
    add("unroll8_R16_th128", [](variant& v) { v.unroll = 8; v.R = 16; v.threads = 128; });
    add("R8_th128", [](variant& v) { v.R = 8; v.threads = 128; });
    add("R8_th64", [](variant& v) { v.R = 8; v.threads = 64; });
    add("R16_th64", [](variant& v) { v.R = 16; v.threads = 64; });
    add("R16_th128_pdl", [](variant& v) { v.R = 16; v.threads = 128; v.pdl = 1; });
    add("R8_th128_pdl", [](variant& v) { v.R = 8; v.threads = 128; v.pdl = 1; });
    add("persistent", [](variant& v) { v.persistent = 1; });
    add("persistent_R8", [](variant& v) { v.persistent = 1; v.R = 8; });
    if (row_bytes % 32 == 0) {
      add("vec32", [](variant& v) { v.vec = 32; });
      add("vec32_pdl", [](variant& v) { v.vec = 32; v.pdl = 1; });
      add("vec32_R8_th128", [](variant& v) { v.vec = 32; v.R = 8; v.threads = 128; });
    }
    if (ct) {
      add("ctrow", [&](variant& v) { v.vn_ct = vn; });
      add("ctrow_unroll8", [&](variant& v) { v.vn_ct = vn; v.unroll = 8; });
      add("ctrow_pdl", [&](variant& v) { v.vn_ct = vn; v.pdl = 1; });
    }
    for (int p : {1, 2, 3, 4}) add(("ldpol" + std::to_string(p)).c_str(), [&](variant& v) { v.policy = p; });
    for (int p : {1, 2, 3, 4}) add(("stpol" + std::to_string(p)).c_str(), [&](variant& v) { v.policy = p << 4; });
    add("bulk4k", [](variant& v) { v.bulk = 1; });
  }
  if (set == "plan") { /* around the shipped plan (128 threads, ~4 KiB batches, 32-byte units, PDL), on bench-sized tables */
    vs.clear();
    for (int R : {2, 4, 8, 16, 32})
      for (int th : {64, 128, 256}) {
        if ((int64_t)R * row_bytes > 32768) continue;
        add(("v32_pdl_R" + std::to_string(R) + "_th" + std::to_string(th)).c_str(), [&](variant& v) { v.vec = 32; v.pdl = 1; v.R = R; v.threads = th; });
      }
    add("v32_pdl_R4_th128_u8", [](variant& v) { v.vec = 32; v.pdl = 1; v.R = 4; v.threads = 128; v.unroll = 8; });
    add("v32_pdl_R8_th128_u2", [](variant& v) { v.vec = 32; v.pdl = 1; v.R = 8; v.threads = 128; v.unroll = 2; });
    add("v16_pdl_R8_th128", [](variant& v) { v.vec = 16; v.pdl = 1; v.R = 8; v.threads = 128; });
  }
  if (set == "link" || set == "all") {
    add("ldpol0", [](variant& v) { v.policy = 0; });
    add("ldpol1_nc", [](variant& v) { v.policy = 1; });
    add("ldpol4_l2_256", [](variant& v) { v.policy = 4; });
    add("unroll8", [](variant& v) { v.unroll = 8; });
    add("unroll8_th512", [](variant& v) { v.unroll = 8; v.threads = 512; });
    add("th1024", [](variant& v) { v.threads = 1024; });
    add("R4", [](variant& v) { v.R = 4; });
    add("R32", [](variant& v) { v.R = 32; });
    add("persistent", [](variant& v) { v.persistent = 1; });
    if (row_bytes % 32 == 0) {
      add("vec32", [](variant& v) { v.vec = 32; });
      add("vec32_ldplain", [](variant& v) { v.vec = 32; v.policy = 2; });
      add("vec32_unroll8", [](variant& v) { v.vec = 32; v.unroll = 8; });
    }
    add("bulk4k", [](variant& v) { v.bulk = 1; });
    add("bulk8k", [](variant& v) { v.bulk = 1; v.slot_kb = 8; });
    add("bulk16k", [](variant& v) { v.bulk = 1; v.slot_kb = 16; });
  }
  return vs;
}

int main(int argc, char** argv)
{
  int64_t row_bytes = 1024, rows = 20000000, n = 1048576;
  int iters = 20, warmup = 5;
  double peak = 6553.0, link_peak = 770.0;
  std::string mode = "local", pattern = "random", set = "small", op = "gather", only;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val      = [&] { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "--row-bytes") row_bytes = atoll(val().c_str());
    else if (a == "--rows") rows = atoll(val().c_str());
    else if (a == "--n") n = atoll(val().c_str());
    else if (a == "--iters") iters = atoi(val().c_str());
    else if (a == "--warmup") warmup = atoi(val().c_str());
    else if (a == "--mode") mode = val();      /* local | uni | bidir */
    else if (a == "--pattern") pattern = val(); /* random | seq */
    else if (a == "--set") set = val();         /* default | small | link | all */
    else if (a == "--op") op = val();           /* gather | scatter */
    else if (a == "--only") only = val();       /* run one variant by name */
    else if (a == "--peak") peak = atof(val().c_str());
    else if (a == "--link-peak") link_peak = atof(val().c_str());
    else {
      fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  const bool gather = op == "gather";
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const bool two = mode != "local";
  if (two && ndev < 2) {
    fprintf(stderr, "mode %s needs 2 GPUs\n", mode.c_str());
    return 3;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  const int nside = two ? 2 : 1;
  side sd[2];
  if (two) {
    for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d));
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, d, 1 - d));
      if (!can) {
        fprintf(stderr, "no peer access %d -> %d\n", d, 1 - d);
        return 3;
      }
      CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    }
  }
  bool have_nvml = two && nvml_open(2);
  std::mt19937_64 rng(1234);
  for (int d = 0; d < nside; ++d) {
    CK(cudaSetDevice(d));
    sd[d].dev = d;
    CK(cudaMalloc(&sd[d].table, (size_t)rows * row_bytes));
    CK(cudaMemset(sd[d].table, d + 1, (size_t)rows * row_bytes));
    CK(cudaMalloc(&sd[d].out, (size_t)n * row_bytes));
    CK(cudaMemset(sd[d].out, 0, (size_t)n * row_bytes));
    CK(cudaStreamCreateWithFlags(&sd[d].s, cudaStreamNonBlocking));
    CK(cudaEventCreate(&sd[d].e0));
    CK(cudaEventCreate(&sd[d].e1));
    std::vector<int64_t> h(n);
    for (int b = 0; b < 8; ++b) {
      if (pattern == "seq") {
        int64_t start = (int64_t)(rng() % (uint64_t)(rows - n));
        for (int64_t i = 0; i < n; ++i) h[i] = start + i;
      } else if (!gather) { /* scatter: distinct rows (a permutation prefix) so no two writers race on one row */
        int64_t stride = rows / n, off = (int64_t)(rng() % (uint64_t)std::max<int64_t>(1, stride));
        for (int64_t i = 0; i < n; ++i) h[i] = i * stride + off;
        std::shuffle(h.begin(), h.end(), rng);
      } else {
        for (int64_t i = 0; i < n; ++i) h[i] = (int64_t)(rng() % (uint64_t)rows);
      }
      CK(cudaMalloc(&sd[d].idx[b], n * sizeof(int64_t)));
      CK(cudaMemcpy(sd[d].idx[b], h.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
  }
  for (int d = 0; d < nside; ++d) { /* the table a device works on: its own (local) or the other device's (uni / bidir) */
    sd[d].tref.mode       = table_ref::FLAT;
    sd[d].tref.nranks     = 1;
    sd[d].tref.has_remote = two ? 1 : 0;
    sd[d].tref.base[0]    = two ? sd[1 - d].table : sd[d].table;
  }
  const int active = mode == "bidir" ? 2 : 1;
  printf("# rowmove_lab op=%s mode=%s pattern=%s row_bytes=%ld rows=%ld n=%ld iters=%d sms=%d nvml=%d\n", op.c_str(), mode.c_str(),
         pattern.c_str(), (long)row_bytes, (long)rows, (long)n, iters, g_sms, (int)have_nvml);
  printf("%-22s %9s %10s %10s %7s", "variant", "ms", "GB/s_out", "GB/s_alg", two ? "f_link" : "f_hbm");
  if (two) printf(" %9s %9s %9s %9s", "d0_rx_dat", "d0_rx_raw", "d0_tx_dat", "d0_tx_raw");
  printf("\n");
  for (const auto& v : variant_set(set, row_bytes)) {
    if (!only.empty() && v.name != only) continue;
    if (v.vec == 32 && row_bytes % 32 != 0) continue;
    for (int it = 0; it < warmup; ++it)
      for (int d = 0; d < active; ++d) {
        CK(cudaSetDevice(d));
        launch(v, gather, sd[d], it % 8, n, row_bytes, two);
      }
    for (int d = 0; d < active; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaStreamSynchronize(sd[d].s));
    }
    nvl_counters c0 = have_nvml ? nvl_read(0) : nvl_counters{};
    for (int d = 0; d < active; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaEventRecord(sd[d].e0, sd[d].s));
    }
    for (int it = 0; it < iters; ++it)
      for (int d = 0; d < active; ++d) {
        CK(cudaSetDevice(d));
        launch(v, gather, sd[d], it % 8, n, row_bytes, two);
      }
    float ms = 0;
    for (int d = 0; d < active; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaEventRecord(sd[d].e1, sd[d].s));
    }
    for (int d = 0; d < active; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaEventSynchronize(sd[d].e1));
      float m = 0;
      CK(cudaEventElapsedTime(&m, sd[d].e0, sd[d].e1));
      ms = std::max(ms, m);
    }
    nvl_counters c1 = have_nvml ? nvl_read(0) : nvl_counters{};
    ms /= iters;
    double out_gbs = (double)n * row_bytes / (ms * 1e-3) / 1e9;
    double alg_gbs = (double)n * (2 * row_bytes + 8) / (ms * 1e-3) / 1e9;
    double frac    = two ? out_gbs / link_peak : alg_gbs / peak;
    printf("%-22s %9.4f %10.1f %10.1f %7.3f", v.name.c_str(), ms, out_gbs, alg_gbs, frac);
    if (two) {
      double t = ms * 1e-3 * iters;
      auto gbs = [&](unsigned long long a, unsigned long long b) { return (double)(b - a) * 1024.0 / t / 1e9; };
      printf(" %9.1f %9.1f %9.1f %9.1f", gbs(c0.data_rx, c1.data_rx), gbs(c0.raw_rx, c1.raw_rx), gbs(c0.data_tx, c1.data_tx),
             gbs(c0.raw_tx, c1.raw_tx));
    }
    printf("\n");
    fflush(stdout);
  }
  /* spot check: device 0's output holds the byte pattern of the table it read */
  if (gather) {
    CK(cudaSetDevice(0));
    std::vector<unsigned char> h(row_bytes);
    CK(cudaMemcpy(h.data(), sd[0].out + (size_t)(n - 1) * row_bytes, row_bytes, cudaMemcpyDeviceToHost));
    unsigned char want = two ? 2 : 1;
    for (auto b : h)
      if (b != want) {
        printf("# CHECK FAILED: output byte %d != %d\n", (int)b, (int)want);
        return 1;
      }
    printf("# check ok\n");
  }
  return 0;
}
