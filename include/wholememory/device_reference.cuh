/*
 * Device-side resolver for wholememory_gref_t (public helper for downstream kernels).
 * Same contract as reference cpp/include/wholememory/device_reference.cuh:25-71:
 * ref[i] yields an lvalue for global element i of a CONTINUOUS (flat) or CHUNKED table.
 *
 * Written for sm_100a: the rank offsets are kept as *byte* offsets and only consulted on the
 * irregular-partition path; the regular path is one 64-bit divide, the flat path is a plain index.
 * The library's own gather/scatter kernels do not use this class (they resolve one index per
 * lane and broadcast, see wholegraph_b200/csrc/gather_scatter.cuh); it is kept for ABI users.
 */
#pragma once
#include <assert.h>
#include <stddef.h>

#include "global_reference.h"

namespace wholememory {

template <typename T>
class device_reference {
 public:
  static constexpr int kMaxRanks = 8; /* same limit as the reference (device_reference.cuh:36) */

  __device__ __forceinline__ explicit device_reference(const wholememory_gref_t& g)
    : base_(g.pointer), chunk_elems_(g.stride / sizeof(T)), nranks_(g.world_size), regular_(g.same_chunk)
  {
    assert(g.stride % sizeof(T) == 0);
    if (chunk_elems_ != 0 && !regular_) {
      assert(nranks_ <= kMaxRanks);
      for (int r = 0; r <= nranks_; ++r) {
        assert(g.rank_memory_offsets[r] % sizeof(T) == 0);
        first_elem_[r] = g.rank_memory_offsets[r] / sizeof(T);
      }
    }
  }
  device_reference() = delete;

  __device__ __forceinline__ T& operator[](size_t i)
  {
    if (chunk_elems_ == 0) return static_cast<T*>(base_)[i];
    T* const* chunks = static_cast<T* const*>(base_);
    if (regular_) {
      size_t owner = i / chunk_elems_;
      return chunks[owner][i - owner * chunk_elems_];
    }
    int owner = 0;
    for (int r = 1; r <= nranks_; ++r) {
      if (i < first_elem_[r]) {
        owner = r - 1;
        break;
      }
    }
    return chunks[owner][i - first_elem_[owner]];
  }

 private:
  void* base_;
  size_t chunk_elems_;
  int nranks_;
  bool regular_;
  size_t first_elem_[kMaxRanks + 1];
};

}  // namespace wholememory
