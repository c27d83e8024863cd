/*
 * Communicators: rank/size bookkeeping, rank-info exchange, peer-capability discovery,
 * split, barrier.  Replaces reference cpp/src/wholememory/communicator.cpp:397-905 for the
 * one-box scope (create_unique_id :397, create_communicator :703-742, exchange_rank_info
 * :526-609, split :744-806, destroy :808-860, support_type_location :358-374).
 *
 * Control plane = wm::bootstrap (AF_UNIX star); NCCL is created lazily and only for the
 * DISTRIBUTED no-peer data plane (nccl_plane.cpp).
 */
#include "wm_internal.hpp"

#include <fcntl.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>

namespace wm {

static std::mutex g_comm_mu;
static std::map<int, wholememory_comm_t> g_comms;
static int g_next_comm_id = 0;

struct rank_info {
  int32_t pid;
  int32_t has_gpu;
  char pci_bus_id[32];
  uint64_t granularity;
};

static void discover_peers(wholememory_comm_t c)
{
  rank_info mine{};
  mine.pid         = (int32_t)getpid();
  mine.has_gpu     = 0;
  mine.granularity = 2u << 20;
  if (cuda_available()) {
    int dev = 0;
    WM_CUDA(cudaGetDevice(&dev));
    c->dev_id    = dev;
    mine.has_gpu = 1;
    WM_CUDA(cudaDeviceGetPCIBusId(mine.pci_bus_id, sizeof(mine.pci_bus_id), dev));
    CUmemAllocationProp prop{};
    prop.type          = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id   = dev;
    size_t g           = 0;
    WM_CUDA(cudaFree(nullptr)); /* make sure the primary context exists before driver calls */
    WM_CU(cu().MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    mine.granularity = g;
  }
  std::vector<rank_info> all(c->world_size);
  c->boot->allgather(&mine, all.data(), sizeof(rank_info));

  c->rank_local_dev.assign(c->world_size, -1);
  size_t gran       = 0;
  bool peers_ok     = mine.has_gpu != 0;
  for (int r = 0; r < c->world_size; ++r) {
    gran = std::max<size_t>(gran, all[r].granularity);
    if (!all[r].has_gpu || !mine.has_gpu) {
      peers_ok = false;
      continue;
    }
    int ldev = -1;
    if (cudaDeviceGetByPCIBusId(&ldev, all[r].pci_bus_id) != cudaSuccess) {
      (void)cudaGetLastError();
      peers_ok = false;
      continue;
    }
    c->rank_local_dev[r] = ldev;
    if (ldev != c->dev_id) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, c->dev_id, ldev) != cudaSuccess) {
        (void)cudaGetLastError();
        can = 0;
      }
      if (!can) peers_ok = false;
    }
  }
  /* two ranks on one GPU is legal (tests), VMM sharing still works */
  c->alloc_granularity = gran;
  /* every rank must agree, otherwise collective allocation would diverge */
  int32_t ok_local = peers_ok ? 1 : 0;
  std::vector<int32_t> ok_all(c->world_size);
  c->boot->allgather(&ok_local, ok_all.data(), sizeof(int32_t));
  c->all_peer_capable = std::all_of(ok_all.begin(), ok_all.end(), [](int32_t v) { return v == 1; });
}

static wholememory_comm_t create_comm(const wholememory_unique_id_t& uid, int rank, int size)
{
  auto* c        = new wholememory_comm_();
  c->world_rank  = rank;
  c->world_size  = size;
  try {
    c->boot = std::make_unique<bootstrap>(uid, rank, size);
    discover_peers(c);
  } catch (...) {
    delete c;
    throw;
  }
  std::lock_guard<std::mutex> lk(g_comm_mu);
  c->comm_id         = g_next_comm_id++;
  g_comms[c->comm_id] = c;
  obj_register(OBJ_COMM, c);
  WM_DEBUG("communicator %d: rank %d/%d dev %d peer_capable=%d granularity=%zu",
           c->comm_id, rank, size, c->dev_id, (int)c->all_peer_capable, c->alloc_granularity);
  return c;
}

static void destroy_comm(wholememory_comm_t c)
{
  obj_unregister(OBJ_COMM, c); /* first: from here on every entry point refuses this communicator */
  {
    std::lock_guard<std::mutex> lk(c->mu);
    /* reference communicator.cpp:808-827: a dying communicator takes its memory with it */
    while (!c->handles.empty()) destroy_handle_locked(c->handles.begin()->second);
  }
  nccl_destroy(c);
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    g_comms.erase(c->comm_id);
  }
  delete c;
}

wholememory_error_code_t destroy_all_communicators_impl()
{
  for (;;) {
    wholememory_comm_t c = nullptr;
    {
      std::lock_guard<std::mutex> lk(g_comm_mu);
      if (g_comms.empty()) break;
      c = g_comms.begin()->second;
    }
    destroy_comm(c);
  }
  return WHOLEMEMORY_SUCCESS;
}

static void fill_unique_id(wholememory_unique_id_t* uid)
{
  memset(uid->internal, 0, sizeof(uid->internal));
  int fd = ::open("/dev/urandom", O_RDONLY | O_CLOEXEC);
  size_t got = 0;
  if (fd >= 0) {
    ssize_t r = ::read(fd, uid->internal, 64);
    got       = r > 0 ? (size_t)r : 0;
    ::close(fd);
  }
  if (got < 16) { /* no urandom: pid + clock + counter is still unique on one box */
    static std::atomic<uint64_t> ctr{0};
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    uint64_t v[3] = {(uint64_t)getpid(), (uint64_t)ts.tv_sec * 1000000000ull + ts.tv_nsec, ctr++};
    memcpy(uid->internal, v, sizeof(v));
  }
  memcpy(uid->internal + 120, "WGB200\0", 8);
}

}  // namespace wm

/* C++-only helper the reference exposes to tests (communicator.hpp:267-288) */
namespace wholememory {
wholememory_error_code_t destroy_all_communicators() noexcept
{
  return wm::guarded("destroy_all_communicators", [] { return wm::destroy_all_communicators_impl(); });
}
}  // namespace wholememory

extern "C" {

wholememory_error_code_t wholememory_create_unique_id(wholememory_unique_id_t* unique_id)
{
  if (unique_id == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  wm::fill_unique_id(unique_id);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_create_communicator(wholememory_comm_t* comm,
                                                         wholememory_unique_id_t unique_id,
                                                         int rank,
                                                         int size)
{
  return wm::guarded("wholememory_create_communicator", [&]() -> wholememory_error_code_t {
    if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
    *comm = wm::create_comm(unique_id, rank, size);
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_split_communicator(wholememory_comm_t* new_comm,
                                                        wholememory_comm_t comm,
                                                        int color,
                                                        int key)
{
  return wm::guarded("wholememory_split_communicator", [&]() -> wholememory_error_code_t {
    if (new_comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
    WM_REQUIRE_LIVE(comm);
    std::unique_lock<std::mutex> lk(comm->mu);
    struct ck {
      int32_t color, key, rank;
    };
    ck mine{color, key, comm->world_rank};
    std::vector<ck> all(comm->world_size);
    comm->boot->allgather(&mine, all.data(), sizeof(ck));
    /* members of my color ordered by (key, old rank) -- reference communicator.cpp:744-806 */
    std::vector<ck> group;
    for (auto& e : all)
      if (e.color == color && color != WHOLEMEMORY_SPILT_NO_COLOR) group.push_back(e);
    std::stable_sort(group.begin(), group.end(), [](const ck& a, const ck& b) {
      return a.key != b.key ? a.key < b.key : a.rank < b.rank;
    });
    /* the first member of every group mints the id; one allgather distributes all of them */
    wholememory_unique_id_t my_uid{};
    if (!group.empty() && group[0].rank == comm->world_rank) wm::fill_unique_id(&my_uid);
    std::vector<wholememory_unique_id_t> uids(comm->world_size);
    comm->boot->allgather(&my_uid, uids.data(), sizeof(my_uid));
    lk.unlock();
    if (group.empty()) {
      *new_comm = nullptr;
      return WHOLEMEMORY_SUCCESS;
    }
    int new_rank = 0;
    for (size_t i = 0; i < group.size(); ++i)
      if (group[i].rank == comm->world_rank) new_rank = (int)i;
    *new_comm = wm::create_comm(uids[group[0].rank], new_rank, (int)group.size());
    (*new_comm)->distributed_backend = comm->distributed_backend;
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_destroy_communicator(wholememory_comm_t comm)
{
  return wm::guarded("wholememory_destroy_communicator", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_LIVE(comm);
    wm::destroy_comm(comm);
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_communicator_support_type_location(
  wholememory_comm_t comm, wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location)
{
  WM_REQUIRE_LIVE(comm);
  if (memory_type != WHOLEMEMORY_MT_CONTINUOUS && memory_type != WHOLEMEMORY_MT_CHUNKED &&
      memory_type != WHOLEMEMORY_MT_DISTRIBUTED)
    return WHOLEMEMORY_NOT_SUPPORTED; /* HIERARCHY: multi-node only */
  if (memory_location == WHOLEMEMORY_ML_HOST) return WHOLEMEMORY_SUCCESS; /* one box: always intranode */
  if (memory_location == WHOLEMEMORY_ML_DEVICE) {
    if (memory_type == WHOLEMEMORY_MT_DISTRIBUTED) return WHOLEMEMORY_SUCCESS;
    return comm->all_peer_capable ? WHOLEMEMORY_SUCCESS : WHOLEMEMORY_NOT_SUPPORTED;
  }
  return WHOLEMEMORY_NOT_SUPPORTED;
}

wholememory_error_code_t wholememory_communicator_get_rank(int* rank, wholememory_comm_t comm)
{
  if (rank == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(comm);
  *rank = comm->world_rank;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_communicator_get_size(int* size, wholememory_comm_t comm)
{
  if (size == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(comm);
  *size = comm->world_size;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_communicator_get_local_size(int* local_size, wholememory_comm_t comm)
{
  if (local_size == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(comm);
  *local_size = comm->world_size; /* single box: every rank is local */
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_communicator_get_clique_info(clique_info_t* clique_info, wholememory_comm_t comm)
{
  if (clique_info == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(comm);
  /* no MNNVL fabric on an HGX box: same answer the reference gives for a zero cluster uuid
   * (communicator.cpp:541-547) */
  *clique_info              = clique_info_t{};
  clique_info->is_in_clique = 0;
  return WHOLEMEMORY_SUCCESS;
}

bool wholememory_communicator_is_bind_to_nvshmem(wholememory_comm_t) { return false; }

wholememory_error_code_t wholememory_communicator_set_distributed_backend(
  wholememory_comm_t comm, wholememory_distributed_backend_t distributed_backend)
{
  WM_REQUIRE_LIVE(comm);
  if (distributed_backend == WHOLEMEMORY_DB_NVSHMEM) {
    WM_ERROR("NVSHMEM backend is not part of this build (no multi-backend dispatch)");
    return WHOLEMEMORY_NOT_SUPPORTED;
  }
  comm->distributed_backend = distributed_backend;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_distributed_backend_t wholememory_communicator_get_distributed_backend(wholememory_comm_t comm)
{
  return wm::live(comm) ? comm->distributed_backend : WHOLEMEMORY_DB_NONE;
}

wholememory_error_code_t wholememory_communicator_barrier(wholememory_comm_t comm)
{
  return wm::guarded("wholememory_communicator_barrier", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_LIVE(comm);
    /* The reference barrier (nccl_comms.cpp:82-86 + sync) only orders its own stream.  Here the
     * rendezvous is on the host, so first drain this rank's device: once every rank has passed
     * the barrier, all peer stores issued before it (scatter into mapped memory) are visible. */
    if (comm->dev_id >= 0) WM_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(comm->mu);
    comm->boot->barrier();
    return WHOLEMEMORY_SUCCESS;
  });
}

bool wholememory_is_intranode_communicator(wholememory_comm_t) { return true; }
bool wholememory_is_intra_mnnvl_communicator(wholememory_comm_t) { return false; }

} /* extern "C" */
