/*
 * Device helpers shared by the unweighted and weighted CSR samplers: the PCG random stream (as RAFT's PCGenerator
 * seeds it -- restated from the published PCG-XSH-RR 64/32 algorithm, parity with RAFT UNPINNED), CSR access through
 * WholeMemory (owner resolve + peer load), and the per-center sample-count kernel.
 */
#pragma once
#include "gather_scatter.cuh"
#include "ops_internal.hpp"

namespace wm {

struct pcg32 {
  uint64_t state, inc;
  __host__ __device__ __forceinline__ uint32_t next_u32()
  {
    uint64_t old = state;
    state        = old * 6364136223846793005ULL + inc;
    uint32_t xs  = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
  }
  __host__ __device__ __forceinline__ void init(uint64_t seed, uint64_t subsequence)
  {
    state = 0;
    inc   = (subsequence << 1u) | 1u;
    next_u32();
    state += seed;
    next_u32();
  }
  __host__ __device__ __forceinline__ int32_t next_positive_int() { return (int32_t)(next_u32() & 0x7fffffffu); }
  __host__ __device__ __forceinline__ int64_t next_positive_int64()
  {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return (int64_t)((lo | (hi << 32)) & 0x7fffffffffffffffULL);
  }
};

struct csr_ref {
  table_ref row_ptr; /* int64 elements */
  table_ref col;     /* int32|int64 elements */
  int64_t row_ptr_offset_bytes;
  int64_t col_offset_bytes;
  /* exchange mode (DISTRIBUTED memory that is not peer-addressable): row_ptr[center] / row_ptr[center+1] were fetched
   * up front into pre_bounds[0..n) / pre_bounds[n..2n), and col_idx is fetched afterwards from the emitted edge ids */
  const int64_t* pre_bounds;
  int have_col;
};

__device__ __forceinline__ int64_t load_row_ptr(const csr_ref& g, int64_t node)
{
  return *reinterpret_cast<const int64_t*>(resolve_table_byte(g.row_ptr, (uint64_t)(g.row_ptr_offset_bytes + node * 8)));
}
__device__ __forceinline__ void node_bounds(const csr_ref& g, int c, int n, int64_t node, int64_t* start, int64_t* end)
{
  if (g.pre_bounds != nullptr) {
    *start = g.pre_bounds[c];
    *end   = g.pre_bounds[n + c];
  } else {
    *start = load_row_ptr(g, node);
    *end   = load_row_ptr(g, node + 1);
  }
}
template <typename ColT>
__device__ __forceinline__ ColT load_col(const csr_ref& g, int64_t edge)
{
  return *reinterpret_cast<const ColT*>(resolve_table_byte(g.col, (uint64_t)(g.col_offset_bytes + edge * (int64_t)sizeof(ColT))));
}

template <typename IdT>
__global__ void sample_count_kernel(csr_ref g, const IdT* __restrict__ centers, int n, int k, int* __restrict__ counts)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n) {
    int64_t node = (int64_t)centers[i], b = 0, e = 0;
    node_bounds(g, i, n, node, &b, &e);
    int deg = (int)(e - b);
    c            = k > 0 ? min(deg, k) : deg;
    if (c < 0) c = 0;
  }
  counts[i] = c; /* counts[n] = 0 so the scan's last slot is the total */
}


}  // namespace wm
