/*
 * Internal definitions shared by the host side of libwholegraph.so (B200-native build).
 * Nothing here is ABI; the ABI is the headers under include/wholememory.
 */
#pragma once

#include <cuda.h>
#include <cuda_runtime_api.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <wholememory/embedding.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/graph_op.h>
#include <wholememory/wholegraph_op.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
#include <wholememory/wholememory_tensor.h>

namespace wm {

/* ------------------------------------------------------------------ logging / errors */
extern int g_log_level; /* LogLevel */
void log_printf(int level, const char* file, int line, const char* fmt, ...)
  __attribute__((format(printf, 4, 5)));
std::string strprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));

#define WM_LOG(level, ...)                                                          \
  do {                                                                              \
    if ((level) <= ::wm::g_log_level) ::wm::log_printf((level), __FILE__, __LINE__, __VA_ARGS__); \
  } while (0)
#define WM_ERROR(...) WM_LOG(LEVEL_ERROR, __VA_ARGS__)
#define WM_WARN(...) WM_LOG(LEVEL_WARN, __VA_ARGS__)
#define WM_INFO(...) WM_LOG(LEVEL_INFO, __VA_ARGS__)
#define WM_DEBUG(...) WM_LOG(LEVEL_DEBUG, __VA_ARGS__)

/* Exception carried up to the C boundary, where it becomes an error code. */
struct error : std::runtime_error {
  wholememory_error_code_t code;
  error(wholememory_error_code_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

#define WM_THROW(code, ...) throw ::wm::error((code), ::wm::strprintf(__VA_ARGS__))
#define WM_EXPECT(cond, code, ...)                 \
  do {                                             \
    if (!(cond)) WM_THROW((code), __VA_ARGS__);    \
  } while (0)
#define WM_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t wm_e_ = (call);                                                                \
    if (wm_e_ != cudaSuccess) {                                                                \
      (void)cudaGetLastError();                                                                \
      WM_THROW(WHOLEMEMORY_CUDA_ERROR, "%s:%d CUDA error %s: %s", __FILE__, __LINE__,          \
               cudaGetErrorName(wm_e_), #call);                                                \
    }                                                                                          \
  } while (0)
#define WM_CU(call)                                                                            \
  do {                                                                                         \
    CUresult wm_r_ = (call);                                                                   \
    if (wm_r_ != CUDA_SUCCESS)                                                                 \
      WM_THROW(WHOLEMEMORY_CUDA_ERROR, "%s:%d CUDA driver error %d (%s): %s", __FILE__,        \
               __LINE__, (int)wm_r_, ::wm::cu_error_string(wm_r_), #call);                     \
  } while (0)

/* Converts anything thrown inside `body` into an error code; nothing crosses the C ABI. */
template <typename F>
wholememory_error_code_t guarded(const char* api, F&& body) noexcept
{
  try {
    return body();
  } catch (const error& e) {
    WM_ERROR("%s: %s", api, e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    WM_ERROR("%s: out of host memory", api);
    return WHOLEMEMORY_OUT_OF_MEMORY;
  } catch (const std::exception& e) {
    WM_ERROR("%s: %s", api, e.what());
    return WHOLEMEMORY_UNKNOW_ERROR;
  } catch (...) {
    WM_ERROR("%s: unknown exception", api);
    return WHOLEMEMORY_UNKNOW_ERROR;
  }
}

inline size_t round_up(size_t v, size_t a) { return a == 0 ? v : (v + a - 1) / a * a; }
inline size_t div_up(size_t v, size_t a) { return (v + a - 1) / a; }

/* ------------------------------------------------------------------ driver / NCCL entry points */
/* The library links neither libcuda nor libnccl: both are resolved at run time so that the .so
 * loads (and its control plane can be unit-tested) on a box with no driver. */
struct cu_api {
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity;
  decltype(&cuMemAddressReserve) MemAddressReserve;
  decltype(&cuMemAddressFree) MemAddressFree;
  decltype(&cuMemCreate) MemCreate;
  decltype(&cuMemRelease) MemRelease;
  decltype(&cuMemMap) MemMap;
  decltype(&cuMemUnmap) MemUnmap;
  decltype(&cuMemSetAccess) MemSetAccess;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle;
  decltype(&cuGetErrorString) GetErrorString;
};
const cu_api& cu(); /* throws WHOLEMEMORY_CUDA_ERROR when no driver is present */
const char* cu_error_string(CUresult r);

bool cuda_available();    /* a usable device exists */
int cuda_device_count();  /* 0 when no driver */
void require_cuda(const char* what); /* throws (loudly) when there is no GPU: no CPU fallback */

/* ------------------------------------------------------------------ control-plane bootstrap */
/* Star-topology host collective over abstract AF_UNIX sockets (single box).  Rank 0 listens on a
 * name derived from the 128-byte unique id.  Carries bytes and file descriptors (SCM_RIGHTS). */
class bootstrap {
 public:
  bootstrap(const wholememory_unique_id_t& uid, int rank, int size);
  ~bootstrap();
  bootstrap(const bootstrap&)            = delete;
  bootstrap& operator=(const bootstrap&) = delete;

  int rank() const { return rank_; }
  int size() const { return size_; }
  /* recv must hold size()*bytes */
  void allgather(const void* send, void* recv, size_t bytes);
  void barrier();
  void broadcast(void* buf, size_t bytes, int root);
  /* send[i] (bytes each) goes to rank i; recv[j] comes from rank j */
  void alltoall(const void* send, void* recv, size_t bytes);
  /* every rank contributes one fd (or -1); returns one fd per rank, owned by the caller
   * (entry for this rank is a dup of my_fd; -1 entries stay -1) */
  std::vector<int> allgather_fds(int my_fd);

 private:
  void send_all(int fd, const void* p, size_t n);
  void recv_all(int fd, void* p, size_t n);
  void send_fd(int sock, int fd);
  int recv_fd(int sock);
  void socket_allgather(const void* send, void* recv, size_t bytes);
  void setup_mailbox();                                              /* collective, over the sockets */
  void mailbox_allgather(const void* send, void* recv, size_t bytes); /* bytes <= mailbox payload */
  void check_sockets_alive();
  struct mailbox;
  int rank_, size_;
  int listen_fd_ = -1;
  std::vector<int> peers_; /* root: socket per rank (index 0 unused); others: peers_[0] = root */
  mailbox* mbox_ = nullptr; /* shared-memory fast path for small collectives; nullptr = sockets only */
};

}  // namespace wm

/* ------------------------------------------------------------------ opaque ABI objects */
struct wholememory_comm_ {
  int world_rank = 0;
  int world_size = 1;
  int comm_id    = 0;
  int dev_id     = -1;  /* CUDA device current when the communicator was created; -1 = none */
  size_t alloc_granularity = 2u << 20;
  wholememory_distributed_backend_t distributed_backend = WHOLEMEMORY_DB_NCCL;
  bool all_peer_capable = false; /* every rank's GPU is visible here and P2P-reachable */
  std::vector<int> rank_local_dev; /* local CUDA ordinal of each rank's GPU, -1 if not visible */
  std::unique_ptr<wm::bootstrap> boot;
  void* nccl_comm = nullptr; /* ncclComm_t, created on first DISTRIBUTED exchange */
  std::mutex mu;
  int next_handle_id = 0;
  std::map<int, wholememory_handle_t> handles;
};

struct wholememory_handle_ {
  int id = 0;
  wholememory_comm_t comm = nullptr;
  wholememory_memory_type_t type         = WHOLEMEMORY_MT_NONE;
  wholememory_memory_location_t location = WHOLEMEMORY_ML_NONE;
  size_t total_size  = 0;
  size_t granularity = 1;

  /* logical partition, bytes */
  std::vector<size_t> part_sizes;   /* world_size   */
  std::vector<size_t> part_offsets; /* world_size+1 */
  size_t chunk_stride = 0;          /* bytes per rank when regular */
  bool regular        = true;       /* owner == offset / chunk_stride */

  /* where things are visible in this process */
  void* flat_base = nullptr;     /* CONTINUOUS: start of the flat range */
  std::vector<void*> rank_base;  /* mapped types (+ peer-mapped DISTRIBUTED): start of rank r's partition */
  void* local_ptr   = nullptr;   /* this rank's partition */
  bool peer_mapped  = false;     /* rank_base[] valid for every rank */

  /* device copies backing the public gref of CHUNKED memory */
  void** d_chunk_table = nullptr;
  size_t* d_offsets    = nullptr;
  wholememory_gref_t gref{};

  /* backing storage */
  enum class backing_t { none, vmm, host_shared, cuda_malloc, pinned_local } backing = backing_t::none;
  CUdeviceptr va          = 0;
  size_t va_size          = 0;
  size_t page_size        = 0;
  std::vector<size_t> map_offsets; /* VA offset of each rank's physical allocation */
  std::vector<size_t> map_sizes;
  std::vector<CUmemGenericAllocationHandle> phys; /* one per rank with map_sizes>0 (0 otherwise) */
  void* host_map   = nullptr; /* mmap of the shared host segment */
  size_t host_size = 0;
  bool host_registered = false;
};

struct wholememory_tensor_ {
  wholememory_handle_t handle = nullptr; /* when is_wm */
  void* storage               = nullptr; /* when !is_wm */
  wholememory_tensor_description_t desc;
  wholememory_tensor_t root = nullptr;
  bool is_wm      = false;
  bool own_handle = false;
};

namespace wm {

/* ------------------------------------------------------------------ live-object registry (registry.cpp) */
enum obj_kind { OBJ_COMM, OBJ_HANDLE, OBJ_TENSOR, OBJ_EMBEDDING, OBJ_OPTIMIZER, OBJ_CACHE_POLICY, OBJ_KINDS };
void obj_register(obj_kind k, const void* p);
void obj_unregister(obj_kind k, const void* p);
bool obj_known(obj_kind k, const void* p);   /* false for nullptr */
std::string api_name(const char* pretty_function); /* "ret f(args)::<lambda()>" -> "f" */
bool live(wholememory_comm_t c);
bool live(wholememory_handle_t h);           /* handle and its communicator both alive */
bool live(wholememory_tensor_t t);           /* tensor alive and, when WholeMemory-backed, its handle too */

/* guard at the top of an extern "C" entry: a null, destroyed or never-issued handle is INVALID_INPUT, never UB */
#define WM_REQUIRE_LIVE(obj)                                                                        \
  do {                                                                                              \
    if (!::wm::live(obj)) {                                                                         \
      WM_ERROR("%s: `%s` is null, was destroyed, or was never issued by this library", ::wm::api_name(__PRETTY_FUNCTION__).c_str(), #obj); \
      return WHOLEMEMORY_INVALID_INPUT;                                                             \
    }                                                                                               \
  } while (0)
#define WM_REQUIRE_KNOWN(kind, obj)                                                                 \
  do {                                                                                              \
    if (!::wm::obj_known((kind), (obj))) {                                                          \
      WM_ERROR("%s: `%s` is null, was destroyed, or was never issued by this library", ::wm::api_name(__PRETTY_FUNCTION__).c_str(), #obj); \
      return WHOLEMEMORY_INVALID_INPUT;                                                             \
    }                                                                                               \
  } while (0)

/* ------------------------------------------------------------------ failing together
 * A rank-local failure inside a collective operation never skips a collective the other ranks are about to enter: local
 * steps run under first_error::attempt (records instead of throwing), the status then travels with fail_together() and
 * every rank throws the first failure. */
struct first_error {
  wholememory_error_code_t code = WHOLEMEMORY_SUCCESS;
  std::string what;
  bool ok() const { return code == WHOLEMEMORY_SUCCESS; }
  template <typename F>
  void attempt(F&& fn)
  {
    if (!ok()) return;
    try {
      fn();
    } catch (const error& e) {
      code = e.code;
      what = e.what();
    } catch (const std::exception& e) {
      code = WHOLEMEMORY_UNKNOW_ERROR;
      what = e.what();
    }
  }
  void rethrow() const
  {
    if (!ok()) throw error(code, what);
  }
};

void fail_together(wholememory_comm_t c, const first_error& mine, const char* stage); /* collective; memory_handle.cpp */

/* ------------------------------------------------------------------ runtime helpers */
wholememory_error_code_t create_handle(wholememory_handle_t* out,
                                       size_t total_size,
                                       wholememory_comm_t comm,
                                       wholememory_memory_type_t type,
                                       wholememory_memory_location_t location,
                                       size_t granularity,
                                       size_t* rank_entry_partition);
void destroy_handle_locked(wholememory_handle_t h); /* comm->mu held */
wholememory_error_code_t destroy_all_communicators_impl();

/* NCCL data plane for DISTRIBUTED memory without peer mapping (exchange.cu / nccl_plane.cpp) */
void nccl_ensure(wholememory_comm_t comm);
void nccl_destroy(wholememory_comm_t comm);
/* byte alltoallv on `stream`: counts/displs in bytes, size world_size */
void nccl_alltoallv_bytes(wholememory_comm_t comm,
                          const void* send,
                          const size_t* send_counts,
                          const size_t* send_displs,
                          void* recv,
                          const size_t* recv_counts,
                          const size_t* recv_displs,
                          cudaStream_t stream);

/* RAII wrapper over the caller's temporary-memory callbacks (protocol: env_func_ptrs.h) */
class temp_buffer {
 public:
  explicit temp_buffer(wholememory_env_func_t* env) : env_(env) {}
  ~temp_buffer() { release(); }
  temp_buffer(const temp_buffer&)            = delete;
  temp_buffer& operator=(const temp_buffer&) = delete;
  void* alloc(size_t elems, wholememory_dtype_t dtype, wholememory_memory_allocation_type_t kind);
  void* device(size_t elems, wholememory_dtype_t dtype) { return alloc(elems, dtype, WHOLEMEMORY_MA_DEVICE); }
  void* host(size_t elems, wholememory_dtype_t dtype) { return alloc(elems, dtype, WHOLEMEMORY_MA_HOST); }
  void* pinned(size_t elems, wholememory_dtype_t dtype) { return alloc(elems, dtype, WHOLEMEMORY_MA_PINNED); }
  void* ptr() const { return ptr_; }
  void release();

 private:
  wholememory_env_func_t* env_;
  void* ctx_ = nullptr;
  void* ptr_ = nullptr;
};

/* variable-size op output allocated into a caller-owned context */
void* output_alloc(wholememory_env_func_t* env,
                   void* memory_context,
                   size_t elems,
                   wholememory_dtype_t dtype,
                   wholememory_memory_allocation_type_t kind = WHOLEMEMORY_MA_DEVICE);

int sm_count(int dev = -1);

/* Small device->host read-back (op output sizes) through a per-thread page-locked staging word: a D2H copy into pageable
 * memory goes through the driver's bounce buffer and costs 10-15 us more than the copy itself.  Copies, then waits for the
 * stream. */
void read_back_sync(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t s);

}  // namespace wm
