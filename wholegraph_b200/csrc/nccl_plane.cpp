/*
 * NCCL data plane, used ONLY for the bucket exchange of DISTRIBUTED memory whose shards are not
 * peer-addressable (and by the gradient exchange).  libnccl is resolved with dlopen so the
 * library has no link-time NCCL dependency; the communicator is created on first use.
 * Replaces the alltoallv primitive of reference cpp/src/wholememory/nccl_comms.cpp:409-437
 * (grouped ncclSend/ncclRecv of bytes).
 */
#include "wm_internal.hpp"

#include <dlfcn.h>
#include <nccl.h>

namespace wm {

namespace {

struct nccl_api {
  decltype(&ncclGetUniqueId) GetUniqueId;
  decltype(&ncclCommInitRank) CommInitRank;
  decltype(&ncclCommDestroy) CommDestroy;
  decltype(&ncclGroupStart) GroupStart;
  decltype(&ncclGroupEnd) GroupEnd;
  decltype(&ncclSend) Send;
  decltype(&ncclRecv) Recv;
  decltype(&ncclGetErrorString) GetErrorString;
};

const nccl_api& nccl()
{
  static nccl_api api{};
  static bool ok = [] {
    /* RTLD_NOLOAD first: reuse the copy torch already loaded (same soname) */
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
    if (!lib) return false;
    bool good = true;
#define WM_SYM(field, name)                                             \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name)); \
  good &= api.field != nullptr;
    WM_SYM(GetUniqueId, "ncclGetUniqueId")
    WM_SYM(CommInitRank, "ncclCommInitRank")
    WM_SYM(CommDestroy, "ncclCommDestroy")
    WM_SYM(GroupStart, "ncclGroupStart")
    WM_SYM(GroupEnd, "ncclGroupEnd")
    WM_SYM(Send, "ncclSend")
    WM_SYM(Recv, "ncclRecv")
    WM_SYM(GetErrorString, "ncclGetErrorString")
#undef WM_SYM
    return good;
  }();
  if (!ok) WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "libnccl.so.2 could not be loaded: %s", dlerror());
  return api;
}

#define WM_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t wm_n_ = (call);                                                                   \
    if (wm_n_ != ncclSuccess)                                                                      \
      WM_THROW(WHOLEMEMORY_COMMUNICATION_ERROR, "%s:%d NCCL error %s: %s", __FILE__, __LINE__,     \
               nccl().GetErrorString(wm_n_), #call);                                               \
  } while (0)

}  // namespace

void nccl_ensure(wholememory_comm_t comm)
{
  std::lock_guard<std::mutex> lk(comm->mu);
  if (comm->nccl_comm != nullptr) return;
  require_cuda("NCCL exchange");
  ncclUniqueId id{};
  if (comm->world_rank == 0) WM_NCCL(nccl().GetUniqueId(&id));
  comm->boot->broadcast(&id, sizeof(id), 0);
  WM_CUDA(cudaSetDevice(comm->dev_id));
  ncclComm_t c = nullptr;
  WM_NCCL(nccl().CommInitRank(&c, comm->world_size, id, comm->world_rank));
  comm->nccl_comm = c;
}

void nccl_destroy(wholememory_comm_t comm)
{
  if (comm->nccl_comm == nullptr) return;
  (void)cudaDeviceSynchronize();
  nccl().CommDestroy(static_cast<ncclComm_t>(comm->nccl_comm));
  comm->nccl_comm = nullptr;
}

void nccl_alltoallv_bytes(wholememory_comm_t comm,
                          const void* send,
                          const size_t* send_counts,
                          const size_t* send_displs,
                          void* recv,
                          const size_t* recv_counts,
                          const size_t* recv_displs,
                          cudaStream_t stream)
{
  nccl_ensure(comm);
  auto c          = static_cast<ncclComm_t>(comm->nccl_comm);
  const auto& api = nccl();
  const int ws = comm->world_size, me = comm->world_rank;
  /* own share never leaves the GPU */
  if (send_counts[me] > 0)
    WM_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + recv_displs[me], static_cast<const char*>(send) + send_displs[me],
                            send_counts[me], cudaMemcpyDeviceToDevice, stream));
  if (ws == 1) return;
  WM_NCCL(api.GroupStart());
  for (int r = 0; r < ws; ++r) {
    if (r == me) continue;
    if (recv_counts[r] > 0) WM_NCCL(api.Recv(static_cast<char*>(recv) + recv_displs[r], recv_counts[r], ncclInt8, r, c, stream));
    if (send_counts[r] > 0)
      WM_NCCL(api.Send(static_cast<const char*>(send) + send_displs[r], send_counts[r], ncclInt8, r, c, stream));
  }
  WM_NCCL(api.GroupEnd());
}

}  // namespace wm
