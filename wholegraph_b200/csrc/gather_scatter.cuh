/*
 * sm_100a row gather / scatter kernels for WholeMemory tables (local HBM, peer HBM over
 * NVLink/NVSwitch through VMM mappings, or registered host memory).
 *
 * Functionally replaces reference cpp/src/wholememory_ops/functions/gather_scatter_func.cuh
 * (gather_func_kernel :253-316, gather_func_sub_warp_kernel :323-376, scatter_func_kernel
 * :519-598) -- same results, different design:
 *
 *  * One kernel body serves both directions ("row_move"): the TABLE side is addressed through an
 *    index + owner lookup, the DENSE side is a plain strided matrix.
 *  * Owner lookup is done ONCE per row by ONE lane (each lane of a warp resolves one index of the
 *    warp's batch: idx -> byte offset -> owning rank -> peer VA), and the resolved pointers are
 *    broadcast with shuffles.  The reference redoes the 64-bit divide in all 32 lanes of every row.
 *    Chunk bases live in kernel parameters (constant bank), not in a device table, so resolving
 *    needs no dependent global load.
 *  * The warp's batch of R rows (~4 KiB) is flattened into R*V vectors of the widest unit the alignment
 *    allows (32 bytes = one 256-bit access, else 16 ... 1) and walked with all 32 lanes, UNROLL vectors
 *    per lane in flight before the first store: bytes-in-flight per SM is what hides HBM (~0.8 us)
 *    and NVSwitch (~2-3 us) latency, there is no idle lane for non-power-of-two rows and no
 *    shared-memory round trip.
 *  * Streaming cache policy: table reads bypass L1 allocation, dense writes are evict-first.
 *
 * Element conversion (table dtype != dense dtype) follows the reference's type_caster chain
 * (gather_scatter_func.cuh:161-208): fp16/bf16 go through float, everything else static_cast.
 */
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "table_ref.hpp"

namespace wm {

__device__ __forceinline__ char* resolve_table_byte(const table_ref& t, uint64_t byte_off)
{
  switch (t.mode) {
    case table_ref::FLAT: return t.base[0] + byte_off;
    case table_ref::CHUNK_REGULAR: {
      uint64_t owner = byte_off / t.chunk_bytes;
      return t.base[owner] + (byte_off - owner * t.chunk_bytes);
    }
    case table_ref::CHUNK_IRREGULAR: {
      int owner = 0;
#pragma unroll 1
      for (int r = 1; r < t.nranks; ++r)
        if (byte_off >= t.first_byte[r]) owner = r;
      return t.base[owner] + (byte_off - t.first_byte[owner]);
    }
    case table_ref::DEVTAB_REGULAR: {
      uint64_t owner = byte_off / t.chunk_bytes;
      return t.dev_bases[owner] + (byte_off - owner * t.chunk_bytes);
    }
    default: {
      int owner = 0;
#pragma unroll 1
      for (int r = 1; r < t.nranks; ++r)
        if (byte_off >= t.dev_first_byte[r]) owner = r;
      return t.dev_bases[owner] + (byte_off - t.dev_first_byte[owner]);
    }
  }
}

/* ---- vector moves with streaming cache hints ---- */
template <int BYTES>
struct vec_t;
/* 32-byte unit: sm_100 has 256-bit global loads/stores (SASS LDG.E.256 / STG.E.256); the default unit wherever every address,
 * stride and the row size are multiples of 32 bytes (+2-4 % over 16-byte units, profiles/README.md round 2). */
struct alignas(32) u32x8 {
  uint32_t v[8];
};
template <>
struct vec_t<32> {
  using type = u32x8;
};
template <>
struct vec_t<16> {
  using type = uint4;
};
template <>
struct vec_t<8> {
  using type = uint2;
};
template <>
struct vec_t<4> {
  using type = uint32_t;
};
template <>
struct vec_t<2> {
  using type = uint16_t;
};
template <>
struct vec_t<1> {
  using type = uint8_t;
};

/* read-once data: do not allocate in L1 */
__device__ __forceinline__ uint4 ld_stream(const uint4* p)
{
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p)
{
  uint2 v;
  asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint16_t ld_stream(const uint16_t* p) { return *p; }
__device__ __forceinline__ uint8_t ld_stream(const uint8_t* p) { return *p; }

__device__ __forceinline__ u32x8 ld_stream(const u32x8* p)
{
  u32x8 r;
  asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
/* L1-allocating form, used when rows can be remote (same choice as the 16-byte path's policy 2) */
__device__ __forceinline__ u32x8 ld_plain(const u32x8* p)
{
  u32x8 r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

/* experiment variants of the 16-byte accesses, selected by row_geom::policy (warp-uniform) */
__device__ __forceinline__ uint4 ld_variant(const uint4* p, int variant)
{
  uint4 v;
  switch (variant) {
    case 1: asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); break;
    case 2: asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); break;
    case 3: asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); break;
    case 4: asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); break;
    default: asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); break;
  }
  return v;
}
__device__ __forceinline__ void st_variant(uint4* p, uint4 v, int variant)
{
  switch (variant) {
    case 1: asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); break;
    case 2: asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); break;
    case 3: asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); break;
    case 4: asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); break;
    default: asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); break;
  }
}

/* write-once data: evict-first so it does not push table lines out of L2 */
__device__ __forceinline__ void st_stream(uint4* p, uint4 v)
{
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(u32x8* p, const u32x8& r)
{
  asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
               "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}
__device__ __forceinline__ void st_plain(u32x8* p, const u32x8& r)
{
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
               "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}
__device__ __forceinline__ void st_stream(uint2* p, uint2 v)
{
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v)
{
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream(uint16_t* p, uint16_t v) { *p = v; }
__device__ __forceinline__ void st_stream(uint8_t* p, uint8_t v) { *p = v; }

__device__ __forceinline__ char* shfl_ptr(char* p, int src_lane)
{
  uint64_t v  = reinterpret_cast<uint64_t>(p);
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src_lane);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src_lane);
  return reinterpret_cast<char*>(((uint64_t)hi << 32) | lo);
}

/*
 * Same-dtype row move = byte copy.  VEC = widest power-of-two vector that divides every address,
 * stride and the row size.  GATHER: dense[i] <- table[idx[i]];  !GATHER: table[idx[i]] <- dense[i].
 *
 * Work split: batches of R = g.batch_rows consecutive indices; warp w takes batches w, w+W, ...
 * Within a batch, lane l < R resolves row l; then the R*V vectors of the batch are walked by all
 * lanes, UNROLL at a time (loads first, then stores).
 */
/* programmatic dependent launch (sm_90+): when the kernel is launched with the programmatic-stream-serialization
 * attribute, the NEXT grid in the stream may start being scheduled as soon as every CTA of this one has passed
 * pdl_launch_dependents(), and this grid's pdl_wait() returns only when the PREVIOUS grid has completed and its writes
 * are visible.  Without the attribute both are no-ops.  Back-to-back gathers overlap their launch ramp with the
 * predecessor's tail this way; nothing that depends on earlier work is read before pdl_wait(). */
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

/* VN_CT > 0: the row is exactly VN_CT vectors (a power of two): the unit -> (row, vector) map is a shift and a mask */
template <typename IdxT, int VEC, bool GATHER, int UNROLL, int VN_CT = 0>
__global__ void __launch_bounds__(1024) row_move_vec_kernel(table_ref tref,
                                                          row_geom g,
                                                          const IdxT* __restrict__ indices,
                                                          int64_t n,
                                                          char* __restrict__ dense)
{
  using V              = typename vec_t<VEC>::type;
  const int lane       = threadIdx.x & 31;
  const int warps_cta  = blockDim.x >> 5;
  const int64_t warp   = (int64_t)blockIdx.x * warps_cta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * warps_cta;
  const int R          = g.batch_rows;
  const int64_t nbatch = (n + R - 1) / R;
  const uint32_t Vn    = VN_CT > 0 ? (uint32_t)VN_CT : (uint32_t)g.units_per_row; /* row size in VEC units */
  const uint64_t magic = g.div_magic;

  pdl_launch_dependents();
  pdl_wait();
  int64_t batch = warp;
  /* software prefetch of the next batch's index */
  IdxT next_idx = -1;
  if (batch < nbatch) {
    int64_t i = batch * R + lane;
    if (lane < R && i < n) next_idx = indices[i];
  }
  for (; batch < nbatch; batch += nwarps) {
    const int64_t first = batch * R;
    const int64_t my_idx = (int64_t)next_idx;
    {
      int64_t nb = batch + nwarps;
      next_idx   = -1;
      if (nb < nbatch) {
        int64_t i = nb * R + lane;
        if (lane < R && i < n) next_idx = indices[i];
      }
    }
    char* trow = nullptr; /* table side of row `lane`; null = skip (negative index / past the end) */
    if (my_idx >= 0)
      trow = resolve_table_byte(tref, (uint64_t)(g.table_offset_bytes + my_idx * g.table_stride_bytes));
    char* dbase = dense + first * g.dense_stride_bytes;
    const uint32_t total = (uint32_t)R * Vn;

    for (uint32_t w0 = 0; w0 < total; w0 += 32u * UNROLL) {
      V val[UNROLL];
      char* dst[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        uint32_t w   = w0 + (uint32_t)u * 32u + (uint32_t)lane;
        /* one-row batches (every row above 4 KiB) need no map at all, so the row length is not limited by the magic's range */
        uint32_t row = VN_CT > 0 ? w / (uint32_t)VN_CT : (R == 1 ? 0u : (uint32_t)(((uint64_t)w * magic) >> 40));
        uint32_t v   = w - row * Vn;
        /* shuffles are executed by all lanes, out-of-range lanes read lane (row & 31) harmlessly */
        char* t   = shfl_ptr(trow, (int)(row & 31u));
        bool live = (w < total) && (t != nullptr);
        char* tp  = t + (size_t)v * VEC;
        char* dp  = dbase + (size_t)row * g.dense_stride_bytes + (size_t)v * VEC;
        dst[u]    = nullptr;
        if (live) {
          if (GATHER) {
            if constexpr (VEC == 16) val[u] = ld_variant(reinterpret_cast<const uint4*>(tp), g.policy & 15);
            else if constexpr (VEC == 32)
              val[u] = (g.policy & 15) == 2 ? ld_plain(reinterpret_cast<const u32x8*>(tp)) : ld_stream(reinterpret_cast<const u32x8*>(tp));
            else val[u] = ld_stream(reinterpret_cast<const V*>(tp));
            dst[u] = dp;
          } else {
            val[u] = ld_stream(reinterpret_cast<const V*>(dp));
            dst[u] = tp;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (dst[u] != nullptr) {
          if (GATHER) {
            if constexpr (VEC == 16) st_variant(reinterpret_cast<uint4*>(dst[u]), val[u], g.policy >> 4);
            else st_stream(reinterpret_cast<V*>(dst[u]), val[u]);
          } else {
            if constexpr (VEC == 32) st_plain(reinterpret_cast<u32x8*>(dst[u]), val[u]);
            else *reinterpret_cast<V*>(dst[u]) = val[u];
          }
        }
      }
    }
  }
}

/* ---- converting move ---- */
template <typename T>
struct cvt_traits {
  using mid = T; /* type the value passes through */
};
template <>
struct cvt_traits<__half> {
  using mid = float;
};
template <>
struct cvt_traits<__nv_bfloat16> {
  using mid = float;
};

/* load as From's intermediate, then narrow/widen to To's intermediate, then store type: the same
 * two-hop chain as the reference's convert_type (double -> half rounds twice, on purpose). */
template <typename From, typename To>
__device__ __forceinline__ To convert_elem(From x)
{
  typename cvt_traits<From>::mid a = static_cast<typename cvt_traits<From>::mid>(x);
  typename cvt_traits<To>::mid b   = static_cast<typename cvt_traits<To>::mid>(a);
  return static_cast<To>(b);
}

template <typename T, int N>
struct alignas(sizeof(T) * N) elem_pack {
  T e[N];
};

/*
 * Converting row move: ALIGN elements per lane per step, loaded/stored as one vector each
 * (host guarantees sizeof(T)*ALIGN <= 16 and alignment on both sides).
 */
template <typename TableT, typename DenseT, typename IdxT, int ALIGN, bool GATHER>
__global__ void __launch_bounds__(256) row_move_cvt_kernel(table_ref tref,
                                                          row_geom g,
                                                          const IdxT* __restrict__ indices,
                                                          int64_t n,
                                                          char* __restrict__ dense)
{
  const int lane       = threadIdx.x & 31;
  const int warps_cta  = blockDim.x >> 5;
  const int64_t warp   = (int64_t)blockIdx.x * warps_cta + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * warps_cta;
  const int R          = g.batch_rows;
  const int64_t nbatch = (n + R - 1) / R;
  const uint32_t Vn    = (uint32_t)g.units_per_row; /* ALIGN-element packs per row */
  const uint64_t magic = g.div_magic;

  for (int64_t batch = warp; batch < nbatch; batch += nwarps) {
    const int64_t first = batch * R;
    int64_t my_idx      = -1;
    if (lane < R && first + lane < n) my_idx = (int64_t)indices[first + lane];
    char* trow = nullptr;
    if (my_idx >= 0)
      trow = resolve_table_byte(tref, (uint64_t)(g.table_offset_bytes + my_idx * g.table_stride_bytes));
    char* dbase          = dense + first * g.dense_stride_bytes;
    const uint32_t total = (uint32_t)R * Vn;
    for (uint32_t w0 = 0; w0 < total; w0 += 32u) {
      uint32_t w   = w0 + (uint32_t)lane;
      uint32_t row = R == 1 ? 0u : (uint32_t)(((uint64_t)w * magic) >> 40);
      uint32_t v   = w - row * Vn;
      char* t      = shfl_ptr(trow, (int)(row & 31u));
      if (w < total && t != nullptr) {
        auto* tp = reinterpret_cast<elem_pack<TableT, ALIGN>*>(t) + v;
        auto* dp = reinterpret_cast<elem_pack<DenseT, ALIGN>*>(dbase + (size_t)row * g.dense_stride_bytes) + v;
        if (GATHER) {
          elem_pack<TableT, ALIGN> in = *tp;
          elem_pack<DenseT, ALIGN> out;
#pragma unroll
          for (int k = 0; k < ALIGN; ++k) out.e[k] = convert_elem<TableT, DenseT>(in.e[k]);
          *dp = out;
        } else {
          elem_pack<DenseT, ALIGN> in = *dp;
          elem_pack<TableT, ALIGN> out;
#pragma unroll
          for (int k = 0; k < ALIGN; ++k) out.e[k] = convert_elem<DenseT, TableT>(in.e[k]);
          *tp = out;
        }
      }
    }
  }
}

}  // namespace wm
