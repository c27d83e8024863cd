/*
 * TMA-bulk row gather (cp.async.bulk, SASS UBLKCP): rows travel global/peer -> shared -> global without
 * touching registers.  Used for same-dtype rows whose size and addresses are multiples of 16 bytes.
 *
 * Each warp owns a ring of STAGES shared-memory slots, each holding R rows.  Lane l of a warp owns row l of
 * every slot: it resolves the index, issues ONE bulk load for its row (all loads of a slot complete on the
 * slot's mbarrier via complete_tx), and once the slot is full issues ONE bulk store of its row to the output
 * and commits it to its own bulk async-group.  A slot is refilled one iteration after its stores were issued
 * (cp.async.bulk.wait_group.read 1), so STAGES-1 slots of loads are always in flight per warp:
 * bytes in flight per SM = warps * (STAGES-1) * R * row_bytes (~96 KiB for 1 KiB rows) with 4 warps per SM.
 * Why it matters on NVLink: the copy engine issues full-line requests and keeps far more bytes in flight per
 * instruction than LDG can, which is what hides the ~2-3 us NVSwitch round trip.
 */
#pragma once
#include "gather_scatter.cuh"

namespace wm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

constexpr int kBulkWarps  = 4;
constexpr int kBulkStages = 4;

/* dynamic smem layout: [warps][stages] mbarriers (8 B each), then [warps][stages][R * row_bytes] row slots */
template <typename IdxT, bool GATHER>
__global__ void __launch_bounds__(kBulkWarps * 32) row_move_bulk_kernel(table_ref tref,
                                                                       row_geom g,
                                                                       const IdxT* __restrict__ indices,
                                                                       int64_t n,
                                                                       char* __restrict__ dense,
                                                                       int row_bytes)
{
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  const int lane        = threadIdx.x & 31;
  const int wid         = threadIdx.x >> 5;
  const int R           = g.batch_rows;
  uint64_t* bars        = reinterpret_cast<uint64_t*>(bulk_smem) + wid * kBulkStages;
  const size_t slot_sz  = (size_t)R * row_bytes;
  unsigned char* slots  = bulk_smem + 128 /* barrier area, kBulkWarps*kBulkStages*8 = 128 B */ + (size_t)wid * kBulkStages * slot_sz;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kBulkStages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int64_t warp   = (int64_t)blockIdx.x * kBulkWarps + wid;
  const int64_t nwarps = (int64_t)gridDim.x * kBulkWarps;
  const int64_t nbatch = (n + R - 1) / R;
  /* number of batches this warp owns: warp, warp + nwarps, ... */
  const int64_t mine = warp < nbatch ? (nbatch - warp + nwarps - 1) / nwarps : 0;

  auto issue = [&](int64_t k) { /* k-th batch of this warp -> slot k % STAGES */
    const int s         = (int)(k % kBulkStages);
    const int64_t first = (warp + k * nwarps) * R;
    const int64_t i     = first + lane;
    char* trow          = nullptr;
    if (lane < R && i < n) {
      const int64_t idx = (int64_t)indices[i];
      if (idx >= 0) trow = resolve_table_byte(tref, (uint64_t)(g.table_offset_bytes + idx * g.table_stride_bytes));
    }
    char* drow          = dense + i * g.dense_stride_bytes;
    const bool live     = trow != nullptr;
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (lane == 0) mbar_expect_tx(bars + s, (uint32_t)__popc(mask) * (uint32_t)row_bytes);
    __syncwarp();
    if (live) bulk_load(slots + (size_t)s * slot_sz + (size_t)lane * row_bytes, GATHER ? trow : drow, (uint32_t)row_bytes, bars + s);
    return live ? (GATHER ? drow : trow) : (char*)nullptr; /* where this lane's row goes */
  };

  /* destination pointers of the in-flight slots live in registers, one per stage */
  char* dst[kBulkStages];
#pragma unroll
  for (int s = 0; s < kBulkStages; ++s) dst[s] = nullptr;
  const int64_t prologue = mine < (kBulkStages - 1) ? mine : (kBulkStages - 1);
  for (int64_t k = 0; k < prologue; ++k) {
    char* d = issue(k);
#pragma unroll
    for (int s = 0; s < kBulkStages; ++s)
      if (s == (int)(k % kBulkStages)) dst[s] = d;
  }
  for (int64_t k = 0; k < mine; ++k) {
    const int s           = (int)(k % kBulkStages);
    const uint32_t parity = (uint32_t)((k / kBulkStages) & 1);
    mbar_wait(bars + s, parity);
    char* d = nullptr;
#pragma unroll
    for (int q = 0; q < kBulkStages; ++q)
      if (q == s) d = dst[q];
    if (d != nullptr) bulk_store(d, slots + (size_t)s * slot_sz + (size_t)lane * row_bytes, (uint32_t)row_bytes);
    bulk_commit();
    /* refill the slot whose stores were issued ONE iteration ago (slot (k-1) % STAGES == (k+STAGES-1) % STAGES) */
    const int64_t kn = k + kBulkStages - 1;
    if (kn < mine) {
      bulk_wait_read<1>();
      __syncwarp();
      char* nd = issue(kn);
#pragma unroll
      for (int q = 0; q < kBulkStages; ++q)
        if (q == (int)(kn % kBulkStages)) dst[q] = nd;
    }
  }
  bulk_wait_read<0>(); /* smem must outlive the last stores' reads */
}

}  // namespace wm
