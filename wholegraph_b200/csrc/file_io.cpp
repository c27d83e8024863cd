/*
 * Partition load / store of WholeMemory from raw binary files (feature loading and checkpointing).
 * Same on-disk format and argument meaning as reference cpp/src/wholememory/file_io.cpp (load_file_to_handle :1860-2057,
 * store_handle_to_file :2059-2165): the file list is ONE logical stream of fixed-size entries; entry e lives in memory at
 * e * memory_entry_stride + memory_offset; every rank loads (stores) exactly the entries of its own partition, staged
 * through a 16 MiB host buffer.  Round-robin sharded files (round_robin_size != 0) are not supported by this build.
 */
#include "wm_internal.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

namespace wm {
namespace {

constexpr size_t kStageBytes = 16u << 20;

wholememory_error_code_t check_layout(wholememory_handle_t h, size_t memory_offset, size_t stride, size_t entry_size)
{
  WM_REQUIRE_LIVE(h);
  if (entry_size == 0 || memory_offset + entry_size > stride) { /* reference file_io.cpp:1868-1874 */
    WM_ERROR("Invalid input, entry_size=%zu, memory_entry_stride=%zu, memory_offset=%zu", entry_size, stride, memory_offset);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (h->granularity % stride != 0) {
    WM_ERROR("Invalid input, memory_entry_stride=%zu, but wm_data_granularity=%zu", stride, h->granularity);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  return WHOLEMEMORY_SUCCESS;
}

struct pinned_stage {
  char* p = nullptr;
  explicit pinned_stage(size_t bytes)
  {
    if (cudaMallocHost(reinterpret_cast<void**>(&p), bytes) != cudaSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
    }
  }
  ~pinned_stage()
  {
    if (p) (void)cudaFreeHost(p);
  }
};

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t wholememory_load_from_file(wholememory_handle_t h,
                                                    size_t memory_offset,
                                                    size_t memory_entry_size,
                                                    size_t file_entry_size,
                                                    const char** file_names,
                                                    int file_count,
                                                    int round_robin_size)
{
  return wm::guarded("wholememory_load_from_file", [&]() -> wholememory_error_code_t {
    using namespace wm;
    const size_t stride = memory_entry_size, esz = file_entry_size;
    auto rc             = check_layout(h, memory_offset, stride, esz);
    if (rc != WHOLEMEMORY_SUCCESS) return rc;
    if (round_robin_size != 0) {
      WM_ERROR("round-robin sharded files are not supported by this build");
      return WHOLEMEMORY_NOT_IMPLEMENTED;
    }
    if (file_names == nullptr || file_count < 0 || file_count >= 65536) {
      WM_ERROR("input file count=%d", file_count);
      return WHOLEMEMORY_INVALID_INPUT;
    }
    /* From here on a failure may be rank-local (a file one rank cannot read, a short read): it is recorded, not thrown,
     * and the closing rendezvous tells every rank -- nobody is left waiting in a barrier for a rank that gave up. */
    first_error err;
    err.attempt([&] {
      std::vector<size_t> first_entry(file_count + 1, 0); /* stream position of each file's first entry */
      for (int i = 0; i < file_count; ++i) {
        struct stat st {};
        if (file_names[i] == nullptr || ::stat(file_names[i], &st) != 0 || access(file_names[i], R_OK) != 0)
          WM_THROW(WHOLEMEMORY_INVALID_INPUT, "input_file[%d] of %d (%s) cannot open for read.", i, file_count, file_names[i] ? file_names[i] : "(null)");
        if ((size_t)st.st_size % esz != 0)
          WM_THROW(WHOLEMEMORY_INVALID_INPUT, "input_file[%d] of %d (%s) size=%zu is not a multiple of entry_size=%zu", i, file_count,
                   file_names[i], (size_t)st.st_size, esz);
        first_entry[i + 1] = first_entry[i] + (size_t)st.st_size / esz;
      }
      const size_t table_entries = h->total_size / stride;
      if (first_entry[file_count] > table_entries)
        WM_THROW(WHOLEMEMORY_INVALID_VALUE, "all %d input files hold %zu entries, but the memory holds only %zu", file_count,
                 first_entry[file_count], table_entries);
      require_cuda("wholememory_load_from_file");
      const int me    = h->comm->world_rank;
      size_t begin    = h->part_offsets[me] / stride;
      size_t end      = std::min((h->part_offsets[me] + h->part_sizes[me]) / stride, first_entry[file_count]);
      char* local     = static_cast<char*>(h->local_ptr);
      const size_t per_stage = std::max<size_t>(1, kStageBytes / esz);
      pinned_stage stage(per_stage * esz);
      WM_EXPECT(stage.p != nullptr, WHOLEMEMORY_OUT_OF_MEMORY, "cannot allocate the %zu-byte staging buffer", per_stage * esz);
      int file = 0;
      for (size_t e = begin; e < end;) {
        while (e >= first_entry[file + 1]) ++file;
        size_t count = std::min({per_stage, end - e, first_entry[file + 1] - e});
        int fd       = ::open(file_names[file], O_RDONLY | O_CLOEXEC);
        WM_EXPECT(fd >= 0, WHOLEMEMORY_SYSTEM_ERROR, "open(%s): %s", file_names[file], strerror(errno));
        size_t want = count * esz, got = 0;
        off_t off   = (off_t)((e - first_entry[file]) * esz);
        while (got < want) {
          ssize_t r = ::pread(fd, stage.p + got, want - got, off + (off_t)got);
          if (r <= 0) {
            ::close(fd);
            WM_THROW(WHOLEMEMORY_SYSTEM_ERROR, "short read from %s", file_names[file]);
          }
          got += (size_t)r;
        }
        ::close(fd);
        char* dst = local + (e - begin) * stride + memory_offset;
        WM_CUDA(cudaMemcpy2D(dst, stride, stage.p, esz, esz, count, cudaMemcpyDefault));
        e += count;
      }
      WM_INFO("rank %d loaded entries [%zu, %zu) from %d file(s)", me, begin, std::max(begin, end), file_count);
    });
    if (!err.ok()) WM_ERROR("%s", err.what.c_str());
    std::lock_guard<std::mutex> lk(h->comm->mu);
    fail_together(h->comm, err, "wholememory_load_from_file"); /* everyone's shard is in place before anyone gathers */
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_store_to_file(wholememory_handle_t h,
                                                   size_t memory_offset,
                                                   size_t memory_entry_stride,
                                                   size_t file_entry_size,
                                                   const char* local_file_name)
{
  return wm::guarded("wholememory_store_to_file", [&]() -> wholememory_error_code_t {
    using namespace wm;
    const size_t stride = memory_entry_stride, esz = file_entry_size;
    auto rc             = check_layout(h, memory_offset, stride, esz);
    if (rc != WHOLEMEMORY_SUCCESS) return rc;
    if (local_file_name == nullptr) return WHOLEMEMORY_INVALID_INPUT;
    require_cuda("wholememory_store_to_file");
    {
      std::lock_guard<std::mutex> lk(h->comm->mu);
      if (h->comm->dev_id >= 0) WM_CUDA(cudaDeviceSynchronize());
      h->comm->boot->barrier(); /* reference :2089: all writers are done before anyone dumps */
    }
    const int me         = h->comm->world_rank;
    const size_t entries = h->part_sizes[me] / stride;
    const char* local    = static_cast<const char*>(h->local_ptr);
    const size_t per_stage = std::max<size_t>(1, kStageBytes / esz);
    pinned_stage stage(per_stage * esz);
    WM_EXPECT(stage.p != nullptr, WHOLEMEMORY_OUT_OF_MEMORY, "cannot allocate the %zu-byte staging buffer", per_stage * esz);
    int fd = ::open(local_file_name, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    WM_EXPECT(fd >= 0, WHOLEMEMORY_SYSTEM_ERROR, "open(%s): %s", local_file_name, strerror(errno));
    for (size_t e = 0; e < entries;) {
      size_t count = std::min(per_stage, entries - e);
      if (cudaMemcpy2D(stage.p, esz, local + e * stride + memory_offset, stride, esz, count, cudaMemcpyDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        ::close(fd);
        WM_THROW(WHOLEMEMORY_CUDA_ERROR, "cudaMemcpy2D failed while storing to %s", local_file_name);
      }
      size_t want = count * esz, put = 0;
      while (put < want) {
        ssize_t w = ::write(fd, stage.p + put, want - put);
        if (w <= 0) {
          ::close(fd);
          WM_THROW(WHOLEMEMORY_SYSTEM_ERROR, "short write to %s: %s", local_file_name, strerror(errno));
        }
        put += (size_t)w;
      }
      e += count;
    }
    ::close(fd);
    return WHOLEMEMORY_SUCCESS;
  });
}

} /* extern "C" */
