/*
 * Process-level runtime: logging, init/finalize, lazy CUDA driver entry points, device props.
 * Replaces reference cpp/src/wholememory/initialize.cpp:35-73, cpp/src/logger.cpp,
 * cpp/src/cuda_macros.cpp, cpp/src/wholememory/system_info.cpp (single-box subset).
 */
#include "wm_internal.hpp"

#include <sys/wait.h>
#include <unistd.h>

namespace wm {

int g_log_level = LEVEL_INFO;

static const char* level_name(int l)
{
  static const char* names[] = {"FATAL", "ERROR", "WARN", "INFO", "DEBUG", "TRACE"};
  return (l >= 0 && l <= LEVEL_TRACE) ? names[l] : "?";
}

void log_printf(int level, const char* file, int line, const char* fmt, ...)
{
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  const char* base = strrchr(file, '/');
  /* same text channel as the reference (stdout, logger.hpp:70-87); errors also to stderr */
  FILE* out = level <= LEVEL_ERROR ? stderr : stdout;
  fprintf(out, "[WM %s] %s:%d %s\n", level_name(level), base ? base + 1 : file, line, msg);
  fflush(out);
}

std::string strprintf(const char* fmt, ...)
{
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  return std::string(msg);
}

/* ---- driver entry points via the (statically linked) runtime ---- */
static std::once_flag g_cu_once;
static cu_api g_cu{};
static bool g_cu_ok = false;

template <typename Fn>
static bool load_entry(const char* name, Fn* out)
{
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st);
  if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || p == nullptr) {
    (void)cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<Fn>(p);
  return true;
}

const cu_api& cu()
{
  std::call_once(g_cu_once, [] {
    bool ok = true;
    ok &= load_entry("cuMemGetAllocationGranularity", &g_cu.MemGetAllocationGranularity);
    ok &= load_entry("cuMemAddressReserve", &g_cu.MemAddressReserve);
    ok &= load_entry("cuMemAddressFree", &g_cu.MemAddressFree);
    ok &= load_entry("cuMemCreate", &g_cu.MemCreate);
    ok &= load_entry("cuMemRelease", &g_cu.MemRelease);
    ok &= load_entry("cuMemMap", &g_cu.MemMap);
    ok &= load_entry("cuMemUnmap", &g_cu.MemUnmap);
    ok &= load_entry("cuMemSetAccess", &g_cu.MemSetAccess);
    ok &= load_entry("cuMemExportToShareableHandle", &g_cu.MemExportToShareableHandle);
    ok &= load_entry("cuMemImportFromShareableHandle", &g_cu.MemImportFromShareableHandle);
    ok &= load_entry("cuGetErrorString", &g_cu.GetErrorString);
    g_cu_ok = ok;
  });
  if (!g_cu_ok)
    WM_THROW(WHOLEMEMORY_CUDA_ERROR, "CUDA driver entry points unavailable (no GPU / driver on this box)");
  return g_cu;
}

const char* cu_error_string(CUresult r)
{
  const char* s = nullptr;
  if (g_cu_ok && g_cu.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown";
}

int cuda_device_count()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

bool cuda_available() { return cuda_device_count() > 0; }

void require_cuda(const char* what)
{
  if (!cuda_available())
    WM_THROW(WHOLEMEMORY_CUDA_ERROR,
             "%s needs a CUDA device and none is usable; this library has no CPU fallback", what);
}

static std::mutex g_prop_mu;
static std::vector<std::unique_ptr<cudaDeviceProp>> g_props;

int sm_count(int dev)
{
  cudaDeviceProp* p = get_device_prop(dev);
  return p ? p->multiProcessorCount : 148;
}

void read_back_sync(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t s)
{
  constexpr size_t kStage = 256;
  thread_local void* stage = nullptr; /* lives as long as the thread; a few hundred bytes of pinned memory */
  if (bytes <= kStage && stage == nullptr) {
    if (cudaMallocHost(&stage, kStage) != cudaSuccess) {
      (void)cudaGetLastError();
      stage = nullptr;
    }
  }
  if (bytes <= kStage && stage != nullptr) {
    WM_CUDA(cudaMemcpyAsync(stage, dev_src, bytes, cudaMemcpyDeviceToHost, s));
    WM_CUDA(cudaStreamSynchronize(s));
    memcpy(host_dst, stage, bytes);
    return;
  }
  WM_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, s));
  WM_CUDA(cudaStreamSynchronize(s));
}

static std::mutex g_init_mu;
static bool g_inited = false;

}  // namespace wm

extern "C" {

cudaDeviceProp* get_device_prop(int dev_id)
{
  if (dev_id < 0) {
    if (cudaGetDevice(&dev_id) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
  }
  std::lock_guard<std::mutex> lk(wm::g_prop_mu);
  if ((int)wm::g_props.size() <= dev_id) wm::g_props.resize(dev_id + 1);
  if (!wm::g_props[dev_id]) {
    auto p = std::make_unique<cudaDeviceProp>();
    if (cudaGetDeviceProperties(p.get(), dev_id) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    wm::g_props[dev_id] = std::move(p);
  }
  return wm::g_props[dev_id].get();
}

wholememory_error_code_t wholememory_init(unsigned int flags, LogLevel log_level)
{
  return wm::guarded("wholememory_init", [&]() -> wholememory_error_code_t {
    std::lock_guard<std::mutex> lk(wm::g_init_mu);
    if (flags != 0) return WHOLEMEMORY_INVALID_INPUT; /* reference initialize.cpp:40 */
    wm::g_log_level = (int)log_level;
    if (wm::g_inited) return WHOLEMEMORY_SUCCESS; /* idempotent; the Python layer re-inits per test */
    wm::g_inited = true;
    int n        = wm::cuda_device_count();
    if (n == 0)
      WM_WARN("wholememory_init: no CUDA device visible; only the host control plane is usable");
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_finalize()
{
  return wm::guarded("wholememory_finalize", [&]() -> wholememory_error_code_t {
    std::lock_guard<std::mutex> lk(wm::g_init_mu);
    wm::g_inited = false;
    return wm::destroy_all_communicators_impl();
  });
}

/* GPU count without creating a CUDA context in the caller (reference parallel_utils.cpp:342) */
int fork_get_device_count()
{
  int fds[2];
  if (pipe(fds) != 0) return -1;
  pid_t pid = fork();
  if (pid < 0) {
    close(fds[0]);
    close(fds[1]);
    return -1;
  }
  if (pid == 0) {
    close(fds[0]);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    ssize_t w = write(fds[1], &n, sizeof(n));
    (void)w;
    close(fds[1]);
    _exit(0);
  }
  close(fds[1]);
  int n        = -1;
  ssize_t got  = read(fds[0], &n, sizeof(n));
  close(fds[0]);
  int status = 0;
  waitpid(pid, &status, 0);
  return got == (ssize_t)sizeof(n) ? n : -1;
}

bool wholememory_is_build_with_nvshmem() { return false; }

} /* extern "C" */
