/*
 * WholeMemory handles: partitioning, allocation and cross-process mapping on one NVSwitch box.
 *
 * Replaces reference cpp/src/wholememory/memory_handle.cpp (partition :1603-1636, allocation
 * strategies :1638-1704, impl classes :312-1218, create/destroy :1793-1990, accessors
 * :1992-2165) and cpp/src/wholememory/wholememory.cpp:112-250.
 *
 * B200-first layout decisions:
 *  - every DEVICE type is backed by CUDA VMM (cuMemCreate + POSIX-fd export + cuMemMap); the
 *    reference's cudaMalloc + cudaIpc path for CHUNKED is gone.  On NVSwitch all peers are
 *    uniform, so CHUNKED = "page-aligned chunk per rank inside ONE VA reservation" and its public
 *    gref is just a device table of chunk starts.
 *  - CONTINUOUS splits 2 MiB pages evenly over ranks so the flat VA is dense (same trick as the
 *    reference's each_rank_multiple_page_strategy); ownership of a boundary row may therefore
 *    differ from the logical partition by < 1 page, which is harmless for a flat VA.
 *  - DISTRIBUTED/DEVICE is ALSO peer-mapped when every rank's GPU is P2P reachable (always true
 *    on an HGX box): the ABI still hides the peers' memory, but gather/scatter use the same
 *    peer-load kernel instead of the NCCL bucket exchange.  WG_DISTRIBUTED_NO_PEER=1 forces the
 *    exchange path (used by the tests and for boxes without full P2P).
 *  - HOST CONTINUOUS/CHUNKED: one memfd segment created by rank 0, fd passed over the bootstrap
 *    socket, mmap'ed and cudaHostRegister'ed by everyone (no SysV keys, no /dev/shm names).
 */
#include "wm_internal.hpp"

#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <numeric>

namespace wm {

namespace {

bool env_flag(const char* name)
{
  const char* v = getenv(name);
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

void make_partition(wholememory_handle_t h, const size_t* rank_entry_partition)
{
  const int ws = h->comm->world_size;
  h->part_sizes.assign(ws, 0);
  h->part_offsets.assign(ws + 1, 0);
  if (rank_entry_partition != nullptr) {
    for (int r = 0; r < ws; ++r) {
      h->part_sizes[r]       = rank_entry_partition[r] * h->granularity;
      h->part_offsets[r + 1] = h->part_offsets[r] + h->part_sizes[r];
    }
    /* "regular" only if owner == offset / stride really holds for every byte.  (The reference
     * derives stride = total/ws for user partitions, memory_handle.cpp:1607, which mis-assigns
     * owners for e.g. {3,3,3,1}; not replicated.) */
    h->chunk_stride = h->part_sizes[0];
    h->regular      = h->chunk_stride > 0;
    for (int r = 0; r < ws && h->regular; ++r) {
      bool last = (r == ws - 1);
      if (h->part_offsets[r] != (size_t)r * h->chunk_stride) h->regular = false;
      if (!last && h->part_sizes[r] != h->chunk_stride) h->regular = false;
      if (last && h->part_sizes[r] > h->chunk_stride) h->regular = false;
    }
    if (!h->regular) h->chunk_stride = *std::max_element(h->part_sizes.begin(), h->part_sizes.end());
    return;
  }
  /* default: ceil(entries / ws) entries per rank, trailing ranks short or empty
   * (reference generate_rank_partition_strategy, memory_handle.cpp:1618-1635) */
  size_t entries  = h->total_size / h->granularity;
  size_t per_rank = div_up(entries, (size_t)ws);
  for (int r = 0; r < ws; ++r) {
    size_t b           = std::min((size_t)r * per_rank, entries);
    size_t e           = std::min((size_t)(r + 1) * per_rank, entries);
    h->part_offsets[r] = b * h->granularity;
    h->part_sizes[r]   = (e - b) * h->granularity;
  }
  h->part_offsets[ws] = entries * h->granularity;
  h->chunk_stride     = per_rank * h->granularity;
  h->regular          = true;
}

/* ---------------- VMM-backed device memory ---------------- */
/* Collective.  Rule for every function that creates backing storage: a rank-local failure never skips a collective the
 * other ranks are about to enter.  Local steps run under `attempt`, which records the first error instead of throwing;
 * the status travels with the next collective (or with the agreement that ends create_handle) and all ranks fail together. */
void vmm_create(wholememory_handle_t h, bool map_peers)
{
  auto* c      = h->comm;
  const int ws = c->world_size, me = c->world_rank;
  h->backing   = wholememory_handle_::backing_t::vmm;
  h->page_size = c->alloc_granularity;
  /* Mapping granule.  Large tables are laid out in 512 MiB-aligned granules (as the reference does for
   * totals >= 16 GiB, memory_handle.cpp:229-230,1686-1687) so the driver can use its largest page size
   * and random row reads miss the TLB less. */
  {
    size_t want = h->total_size >= ((size_t)16 << 30) ? (size_t)512 << 20 : 0;
    if (want > h->page_size) h->page_size = round_up(want, c->alloc_granularity);
  }
  h->map_offsets.assign(ws, 0);
  h->map_sizes.assign(ws, 0);
  h->phys.assign(ws, 0);
  const size_t page = h->page_size;
  if (h->type == WHOLEMEMORY_MT_CONTINUOUS) {
    size_t pages = div_up(h->total_size, page);
    for (int r = 0; r < ws; ++r) {
      size_t b          = (size_t)r * pages / ws, e = (size_t)(r + 1) * pages / ws;
      h->map_offsets[r] = b * page;
      h->map_sizes[r]   = (e - b) * page;
    }
    h->va_size = pages * page;
  } else {
    size_t off = 0;
    for (int r = 0; r < ws; ++r) {
      h->map_offsets[r] = off;
      h->map_sizes[r]   = round_up(h->part_sizes[r], page);
      off += h->map_sizes[r];
    }
    h->va_size = off;
  }
  h->rank_base.assign(ws, nullptr);
  const bool share = map_peers && ws > 1;

  /* phase A (local): device, VA range, my physical shard */
  first_error err;
  err.attempt([&] {
    require_cuda("device WholeMemory allocation");
    WM_CUDA(cudaSetDevice(c->dev_id));
    if (h->va_size == 0) return;
    const auto& d = cu();
    WM_CU(d.MemAddressReserve(&h->va, h->va_size, page, 0, 0));
    CUmemAllocationProp prop{};
    prop.type                 = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type        = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id          = c->dev_id;
    prop.requestedHandleTypes = share ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
    if (h->map_sizes[me] > 0) {
      CUresult created = d.MemCreate(&h->phys[me], h->map_sizes[me], &prop, 0);
      if (created != CUDA_SUCCESS) {
        h->phys[me] = 0;
        if (created == CUDA_ERROR_OUT_OF_MEMORY)
          WM_THROW(WHOLEMEMORY_OUT_OF_MEMORY, "cuMemCreate(%zu bytes) out of device memory on rank %d", h->map_sizes[me], me);
        WM_THROW(WHOLEMEMORY_CUDA_ERROR, "cuMemCreate(%zu bytes) failed on rank %d: CUDA driver error %d (%s)", h->map_sizes[me], me,
                 (int)created, cu_error_string(created));
      }
    }
  });
  fail_together(c, err, "device memory allocation"); /* collective 1 */
  if (h->va_size == 0) return;
  const auto& d = cu();

  /* phase B: my shard's fd to everyone (collective 2 runs even when my export failed: fd -1) */
  if (share) {
    int my_fd = -1;
    err.attempt([&] {
      if (h->phys[me] != 0)
        WM_CU(d.MemExportToShareableHandle(&my_fd, h->phys[me], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    });
    std::vector<int> fds = c->boot->allgather_fds(my_fd);
    if (my_fd >= 0) ::close(my_fd);
    for (int r = 0; r < ws; ++r) {
      if (r != me && h->map_sizes[r] > 0) {
        if (fds[r] < 0) err.attempt([&] { WM_THROW(WHOLEMEMORY_CUDA_ERROR, "rank %d could not export its shard", r); });
        else
          err.attempt([&] {
            WM_CU(d.MemImportFromShareableHandle(&h->phys[r], (void*)(uintptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
          });
      }
      if (fds[r] >= 0) ::close(fds[r]);
    }
  }
  /* phase C (local): map and open every shard; a failure here is reported by the agreement that ends create_handle */
  err.attempt([&] {
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id   = c->dev_id;
    acc.flags         = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (int r = 0; r < ws; ++r) {
      if (h->phys[r] == 0) continue;
      WM_CU(d.MemMap(h->va + h->map_offsets[r], h->map_sizes[r], 0, h->phys[r], 0));
      WM_CU(d.MemSetAccess(h->va + h->map_offsets[r], h->map_sizes[r], &acc, 1));
    }
  });
  err.rethrow();
  char* base = reinterpret_cast<char*>(h->va);
  if (h->type == WHOLEMEMORY_MT_CONTINUOUS) {
    h->flat_base = base;
    for (int r = 0; r < ws; ++r) h->rank_base[r] = base + h->part_offsets[r];
    h->peer_mapped = true;
  } else {
    for (int r = 0; r < ws; ++r) h->rank_base[r] = (share || r == me) ? base + h->map_offsets[r] : nullptr;
    h->peer_mapped = share || ws == 1;
  }
  h->local_ptr = h->part_sizes[me] > 0 ? h->rank_base[me] : nullptr;
}

void vmm_destroy(wholememory_handle_t h) noexcept
{
  if (h->va == 0) return;
  const auto& d = cu();
  (void)cudaDeviceSynchronize(); /* no kernel may still be reading the range */
  for (size_t r = 0; r < h->phys.size(); ++r) {
    if (h->phys[r] == 0) continue;
    d.MemUnmap(h->va + h->map_offsets[r], h->map_sizes[r]);
    d.MemRelease(h->phys[r]);
  }
  d.MemAddressFree(h->va, h->va_size);
  h->va = 0;
}

/* ---------------- shared host segment ---------------- */
void host_shared_create(wholememory_handle_t h)
{
  auto* c    = h->comm;
  h->backing = wholememory_handle_::backing_t::host_shared;
  h->rank_base.assign(c->world_size, nullptr);
  if (h->total_size == 0) return;
  /* rank 0's failure to create the segment travels as fd -1: every rank still runs the fd exchange and all of them
   * fail on "no shared segment"; later rank-local failures (mmap, cudaHostRegister) reach the others through the
   * agreement that ends create_handle */
  int fd = -1;
  first_error err;
  err.attempt([&] { require_cuda("host WholeMemory allocation (cudaHostRegister)"); });
  if (c->world_rank == 0 && err.ok()) {
    fd = memfd_create("wgb200_host_wm", MFD_CLOEXEC);
    if (fd < 0) {
      WM_ERROR("memfd_create: %s", strerror(errno));
    } else if (ftruncate(fd, (off_t)h->total_size) != 0) {
      WM_ERROR("ftruncate(%zu): %s", h->total_size, strerror(errno));
      ::close(fd);
      fd = -1;
    }
  }
  std::vector<int> fds = c->boot->allgather_fds(fd);
  if (fd >= 0) ::close(fd);
  int seg = fds[0];
  for (size_t r = 1; r < fds.size(); ++r)
    if (fds[r] >= 0) ::close(fds[r]);
  err.rethrow();
  WM_EXPECT(seg >= 0, WHOLEMEMORY_SYSTEM_ERROR, "no shared segment fd received (rank 0 could not create it)");
  void* p = mmap(nullptr, h->total_size, PROT_READ | PROT_WRITE, MAP_SHARED, seg, 0);
  ::close(seg);
  WM_EXPECT(p != MAP_FAILED, WHOLEMEMORY_SYSTEM_ERROR, "mmap(%zu): %s", h->total_size, strerror(errno));
  h->host_map  = p;
  h->host_size = h->total_size;
  WM_CUDA(cudaHostRegister(p, h->total_size, cudaHostRegisterPortable | cudaHostRegisterMapped));
  h->host_registered = true;
  void* dptr         = nullptr;
  WM_CUDA(cudaHostGetDevicePointer(&dptr, p, 0));
  WM_EXPECT(dptr == p, WHOLEMEMORY_NOT_SUPPORTED, "registered host memory is not identity-mapped on this device");
  h->flat_base = p;
  char* base   = static_cast<char*>(p);
  for (int r = 0; r < c->world_size; ++r) h->rank_base[r] = base + h->part_offsets[r];
  h->peer_mapped = true;
  h->local_ptr   = h->part_sizes[c->world_rank] > 0 ? h->rank_base[c->world_rank] : nullptr;
  /* every rank clears its own slice (reference memory_handle.cpp:571) */
  if (h->local_ptr) memset(h->local_ptr, 0, h->part_sizes[c->world_rank]);
}

void host_shared_destroy(wholememory_handle_t h) noexcept
{
  if (h->host_map == nullptr) return;
  (void)cudaDeviceSynchronize();
  if (h->host_registered) (void)cudaHostUnregister(h->host_map);
  munmap(h->host_map, h->host_size);
  h->host_map = nullptr;
}

void pinned_local_create(wholememory_handle_t h)
{
  require_cuda("pinned host WholeMemory allocation");
  auto* c    = h->comm;
  h->backing = wholememory_handle_::backing_t::pinned_local;
  h->rank_base.assign(c->world_size, nullptr);
  size_t bytes = h->part_sizes[c->world_rank];
  if (bytes == 0) return;
  void* p = nullptr;
  WM_CUDA(cudaMallocHost(&p, bytes));
  h->local_ptr                 = p;
  h->rank_base[c->world_rank] = p;
  h->peer_mapped               = c->world_size == 1;
}

void build_public_gref(wholememory_handle_t h)
{
  auto* c              = h->comm;
  h->gref              = wholememory_gref_t{};
  h->gref.world_size   = c->world_size;
  h->gref.same_chunk   = true;
  if (h->type == WHOLEMEMORY_MT_DISTRIBUTED) return; /* nothing public */
  if (h->flat_base != nullptr && (h->type == WHOLEMEMORY_MT_CONTINUOUS || h->location == WHOLEMEMORY_ML_HOST)) {
    h->gref.pointer = h->flat_base; /* flat: stride 0 (reference memory_handle.cpp:443-450, :676-683) */
    h->gref.stride  = 0;
    return;
  }
  if (h->type == WHOLEMEMORY_MT_CHUNKED && h->va_size > 0) {
    /* device tables, as the ABI promises (reference memory_handle.cpp:1175-1187) */
    const int ws = c->world_size;
    WM_CUDA(cudaMalloc((void**)&h->d_chunk_table, sizeof(void*) * ws));
    WM_CUDA(cudaMemcpy(h->d_chunk_table, h->rank_base.data(), sizeof(void*) * ws, cudaMemcpyHostToDevice));
    WM_CUDA(cudaMalloc((void**)&h->d_offsets, sizeof(size_t) * (ws + 1)));
    WM_CUDA(cudaMemcpy(h->d_offsets, h->part_offsets.data(), sizeof(size_t) * (ws + 1), cudaMemcpyHostToDevice));
    h->gref.pointer             = h->d_chunk_table;
    h->gref.rank_memory_offsets = h->d_offsets;
    h->gref.stride              = h->chunk_stride;
    h->gref.same_chunk          = h->regular;
  }
}

void release_storage(wholememory_handle_t h) noexcept
{
  if (h->d_chunk_table) (void)cudaFree(h->d_chunk_table);
  if (h->d_offsets) (void)cudaFree(h->d_offsets);
  h->d_chunk_table = nullptr;
  h->d_offsets     = nullptr;
  switch (h->backing) {
    case wholememory_handle_::backing_t::vmm: vmm_destroy(h); break;
    case wholememory_handle_::backing_t::host_shared: host_shared_destroy(h); break;
    case wholememory_handle_::backing_t::pinned_local:
      if (h->local_ptr) {
        (void)cudaDeviceSynchronize();
        (void)cudaFreeHost(h->local_ptr);
      }
      break;
    default: break;
  }
  h->backing = wholememory_handle_::backing_t::none;
}

}  // namespace

/* every rank learns every rank's status; all throw the first failure (lowest rank) */
void fail_together(wholememory_comm_t c, const first_error& mine, const char* stage)
{
  if (c->world_size == 1) {
    mine.rethrow();
    return;
  }
  int32_t code = (int32_t)mine.code;
  std::vector<int32_t> all(c->world_size, 0);
  c->boot->allgather(&code, all.data(), sizeof(code));
  mine.rethrow(); /* my own failure carries the detailed message */
  for (int r = 0; r < c->world_size; ++r)
    if (all[r] != (int32_t)WHOLEMEMORY_SUCCESS)
      WM_THROW((wholememory_error_code_t)all[r], "%s failed on rank %d (see that rank's log)", stage, r);
}

wholememory_error_code_t create_handle(wholememory_handle_t* out,
                                       size_t total_size,
                                       wholememory_comm_t comm,
                                       wholememory_memory_type_t type,
                                       wholememory_memory_location_t location,
                                       size_t granularity,
                                       size_t* rank_entry_partition)
{
  if (out == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  *out = nullptr;
  WM_REQUIRE_LIVE(comm);
  if (granularity == 0 || total_size % granularity != 0) return WHOLEMEMORY_INVALID_VALUE;
  if (type != WHOLEMEMORY_MT_CONTINUOUS && type != WHOLEMEMORY_MT_CHUNKED && type != WHOLEMEMORY_MT_DISTRIBUTED) {
    WM_ERROR("memory type %d is not supported by this build (HIERARCHY is multi-node only)", (int)type);
    return WHOLEMEMORY_NOT_SUPPORTED;
  }
  if (location != WHOLEMEMORY_ML_DEVICE && location != WHOLEMEMORY_ML_HOST) return WHOLEMEMORY_INVALID_INPUT;
  if (rank_entry_partition != nullptr) {
    size_t entries = 0;
    for (int r = 0; r < comm->world_size; ++r) {
      if ((int64_t)rank_entry_partition[r] <= 0) return WHOLEMEMORY_INVALID_VALUE; /* reference :1807 */
      entries += rank_entry_partition[r];
    }
    if (entries * granularity != total_size) {
      WM_ERROR("partition entries * granularity (%zu*%zu) != total size (%zu)", entries, granularity, total_size);
      return WHOLEMEMORY_INVALID_VALUE;
    }
  }
  if (wholememory_communicator_support_type_location(comm, type, location) != WHOLEMEMORY_SUCCESS) {
    WM_ERROR("communicator does not support memory type %d at location %d", (int)type, (int)location);
    return WHOLEMEMORY_NOT_SUPPORTED;
  }

  std::lock_guard<std::mutex> lk(comm->mu);
  /* collective sanity check (the reference's WM_COMM_CHECK_ALL_SAME, communicator.hpp:234-263) */
  /* every entry of a custom partition must agree too (the reference checks each one); an FNV-1a hash of the array
   * travels with the scalars, so ranks with different partitions of the same total fail here instead of building
   * different owner tables */
  uint64_t part_hash = 0;
  if (rank_entry_partition != nullptr) {
    part_hash = 1469598103934665603ull;
    for (int r = 0; r < comm->world_size; ++r)
      for (int b = 0; b < 8; ++b) part_hash = (part_hash ^ ((rank_entry_partition[r] >> (8 * b)) & 0xff)) * 1099511628211ull;
  }
  struct params {
    uint64_t total, gran, part_hash;
    int32_t type, location, id, custom;
  } mine{total_size, granularity, part_hash, (int32_t)type, (int32_t)location, comm->next_handle_id, rank_entry_partition != nullptr};
  std::vector<params> all(comm->world_size);
  comm->boot->allgather(&mine, all.data(), sizeof(params));
  for (auto& p : all)
    if (memcmp(&p, &mine, sizeof(params)) != 0)
      WM_THROW(WHOLEMEMORY_LOGIC_ERROR, "wholememory_malloc called with different arguments on different ranks");

  auto h         = std::make_unique<wholememory_handle_>();
  h->id          = comm->next_handle_id++;
  h->comm        = comm;
  h->type        = type;
  h->location    = location;
  h->total_size  = total_size;
  h->granularity = granularity;
  make_partition(h.get(), rank_entry_partition);
  first_error err;
  err.attempt([&] {
    if (location == WHOLEMEMORY_ML_DEVICE) {
      bool map_peers = true;
      if (type == WHOLEMEMORY_MT_DISTRIBUTED)
        map_peers = comm->all_peer_capable && !env_flag("WG_DISTRIBUTED_NO_PEER");
      vmm_create(h.get(), map_peers);
    } else if (type == WHOLEMEMORY_MT_DISTRIBUTED) {
      pinned_local_create(h.get());
    } else {
      host_shared_create(h.get());
    }
    build_public_gref(h.get());
  });
  /* Everyone mapped before anyone touches a peer -- and everyone learns whether everyone succeeded: a rank whose import,
   * map or registration failed takes all ranks out together instead of leaving them with a handle it does not have.
   * (Failures that vmm_create / host_shared_create already made collective arrive here on every rank at once.) */
  try {
    fail_together(comm, err, "WholeMemory allocation");
  } catch (...) {
    release_storage(h.get());
    throw;
  }
  comm->handles[h->id] = h.get();
  obj_register(OBJ_HANDLE, h.get());
  *out                 = h.release();
  return WHOLEMEMORY_SUCCESS;
}

void destroy_handle_locked(wholememory_handle_t h)
{
  auto* comm = h->comm;
  obj_unregister(OBJ_HANDLE, h);
  if (comm->dev_id >= 0) (void)cudaDeviceSynchronize();
  comm->boot->barrier(); /* nobody still reads my shard */
  release_storage(h);
  comm->boot->barrier();
  comm->handles.erase(h->id);
  delete h;
}

}  // namespace wm

extern "C" {

wholememory_error_code_t wholememory_malloc(wholememory_handle_t* wholememory_handle_ptr,
                                            size_t total_size,
                                            wholememory_comm_t comm,
                                            wholememory_memory_type_t memory_type,
                                            wholememory_memory_location_t memory_location,
                                            size_t data_granularity,
                                            size_t* rank_entry_partition)
{
  return wm::guarded("wholememory_malloc", [&] {
    return wm::create_handle(wholememory_handle_ptr, total_size, comm, memory_type, memory_location,
                             data_granularity, rank_entry_partition);
  });
}

wholememory_error_code_t wholememory_free(wholememory_handle_t h)
{
  return wm::guarded("wholememory_free", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_LIVE(h);
    auto* comm = h->comm;
    std::lock_guard<std::mutex> lk(comm->mu);
    if (comm->handles.find(h->id) == comm->handles.end()) return WHOLEMEMORY_INVALID_VALUE;
    wm::destroy_handle_locked(h);
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_get_communicator(wholememory_comm_t* comm, wholememory_handle_t h)
{
  if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  *comm = h->comm;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_get_local_communicator(wholememory_comm_t* comm, wholememory_handle_t h)
{
  if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  return WHOLEMEMORY_NOT_SUPPORTED; /* HIERARCHY only (reference memory_handle.cpp:1999-2001) */
}

wholememory_error_code_t wholememory_get_cross_communicator(wholememory_comm_t* comm, wholememory_handle_t h)
{
  if (comm == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  return WHOLEMEMORY_NOT_SUPPORTED;
}

/* value-returning getters have no error channel: a dead handle reads as "none" / 0 */
wholememory_memory_type_t wholememory_get_memory_type(wholememory_handle_t h) { return wm::live(h) ? h->type : WHOLEMEMORY_MT_NONE; }
wholememory_memory_location_t wholememory_get_memory_location(wholememory_handle_t h)
{
  return wm::live(h) ? h->location : WHOLEMEMORY_ML_NONE;
}
wholememory_distributed_backend_t wholememory_get_distributed_backend(wholememory_handle_t h)
{
  return wm::live(h) ? h->comm->distributed_backend : WHOLEMEMORY_DB_NONE;
}
size_t wholememory_get_total_size(wholememory_handle_t h) { return wm::live(h) ? h->total_size : 0; }
size_t wholememory_get_data_granularity(wholememory_handle_t h) { return wm::live(h) ? h->granularity : 0; }

wholememory_error_code_t wholememory_get_local_memory(void** local_ptr,
                                                      size_t* local_size,
                                                      size_t* local_offset,
                                                      wholememory_handle_t h)
{
  WM_REQUIRE_LIVE(h);
  int me = h->comm->world_rank;
  if (local_ptr) *local_ptr = h->local_ptr;
  if (local_size) *local_size = h->part_sizes[me];
  if (local_offset) *local_offset = h->part_offsets[me];
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_get_local_size(size_t* local_size, wholememory_handle_t h)
{
  return wholememory_get_local_memory(nullptr, local_size, nullptr, h);
}

wholememory_error_code_t wholememory_get_local_offset(size_t* local_offset, wholememory_handle_t h)
{
  return wholememory_get_local_memory(nullptr, nullptr, local_offset, h);
}

wholememory_error_code_t wholememory_get_rank_memory(void** rank_memory_ptr,
                                                     size_t* rank_memory_size,
                                                     size_t* rank_memory_offset,
                                                     int rank,
                                                     wholememory_handle_t h)
{
  WM_REQUIRE_LIVE(h);
  if (rank < 0 || rank >= h->comm->world_size) return WHOLEMEMORY_INVALID_INPUT;
  /* DISTRIBUTED memory is private by contract even when this build peer-maps it internally */
  if (h->type == WHOLEMEMORY_MT_DISTRIBUTED) {
    if (rank_memory_ptr) *rank_memory_ptr = nullptr;
    if (rank_memory_size) *rank_memory_size = 0;
    if (rank_memory_offset) *rank_memory_offset = 0;
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (rank_memory_ptr) *rank_memory_ptr = h->rank_base[rank];
  if (rank_memory_size) *rank_memory_size = h->part_sizes[rank];
  if (rank_memory_offset) *rank_memory_offset = h->part_offsets[rank];
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_equal_entry_partition_plan(size_t* entry_per_rank,
                                                                size_t total_entry_count,
                                                                int world_size)
{
  if (entry_per_rank == nullptr || world_size <= 0) return WHOLEMEMORY_INVALID_INPUT;
  *entry_per_rank = wm::div_up(total_entry_count, (size_t)world_size);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_get_global_pointer(void** global_ptr, wholememory_handle_t h)
{
  if (global_ptr == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  bool flat   = h->type == WHOLEMEMORY_MT_CONTINUOUS || (h->type == WHOLEMEMORY_MT_CHUNKED && h->location == WHOLEMEMORY_ML_HOST);
  *global_ptr = flat ? h->flat_base : nullptr;
  return *global_ptr ? WHOLEMEMORY_SUCCESS : WHOLEMEMORY_INVALID_INPUT;
}

wholememory_error_code_t wholememory_get_global_reference(wholememory_gref_t* gref, wholememory_handle_t h)
{
  if (gref == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  *gref = h->gref;
  return gref->pointer ? WHOLEMEMORY_SUCCESS : WHOLEMEMORY_INVALID_INPUT;
}

wholememory_error_code_t wholememory_get_rank_partition_sizes(size_t* rank_mem_sizes, wholememory_handle_t h)
{
  if (rank_mem_sizes == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  std::copy(h->part_sizes.begin(), h->part_sizes.end(), rank_mem_sizes);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_get_rank_partition_offsets(size_t* rank_mem_offsets, wholememory_handle_t h)
{
  if (rank_mem_offsets == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(h);
  std::copy(h->part_offsets.begin(), h->part_offsets.end(), rank_mem_offsets);
  return WHOLEMEMORY_SUCCESS;
}

} /* extern "C" */
