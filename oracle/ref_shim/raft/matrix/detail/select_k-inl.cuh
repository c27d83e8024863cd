/*
 * Stand-in for rapidsai/raft branch-24.12 raft/matrix/detail/select_k-inl.cuh (RAFT is not vendored here).
 *
 * The reference's sparse-optimizer translation unit (cpp/src/wholememory_ops/functions/embedding_optimizer_func.cu:22)
 * includes embedding_cache_func.cuh only for `CacheLineInfo`; that header also names RAFT's warp-level top-k queue
 * inside the class template `CacheSetUpdater<NodeIDT>` (embedding_cache_func.cuh:163-164, :256-260, :350-374), which
 * the optimizer TU never instantiates.  Two-phase lookup still needs the names to exist, so this file DECLARES them.
 * The members are deliberately left undefined: any TU that instantiated the cache updater would fail to link, i.e.
 * this shim can only ever let the optimizer kernels build, never fake the embedding cache.
 * Test infrastructure (oracle/_ref build); not part of the product.
 */
#pragma once

namespace raft {

__device__ int laneId(); /* declared only */

namespace matrix {
namespace detail {
namespace select {
namespace warpsort {

template <int Capacity, bool Ascending, typename T, typename IdxT>
class warp_sort_immediate {
 public:
  __device__ explicit warp_sort_immediate(int k);
  __device__ void add(T val, IdxT idx);
  __device__ void done();
  __device__ void store(T* out, IdxT* out_idx) const;
};

}  // namespace warpsort
}  // namespace select
}  // namespace detail
}  // namespace matrix
}  // namespace raft
