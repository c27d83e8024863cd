/*
 * Stand-in for rapidsai/raft branch-24.12 raft/random/detail/rng_device.cuh -- ONLY what the reference's unweighted
 * sampler and its host replay use: DeviceState<PCGenerator>, PCGenerator(state, subsequence), next(T&) for
 * int32/int64/uint32/uint64/float/double, UniformDistParams<T> and custom_next.
 *
 * RAFT itself is an un-vendored dependency, so this is a RESTATEMENT of the published PCG-XSH-RR 64/32 generator with
 * the seeding RAFT's PCGenerator applies, identical to oracle/wm_oracle.c:412-441 and to
 * wholegraph_b200/csrc/sample_common.cuh.  It is NOT verified against RAFT's source.  What it buys: the reference's own
 * sampling kernels (count -> scan -> BlockRadixSort + pointer jumping) compile unmodified and run with the same random
 * stream as this repo's sampler, so the SELECTION algorithm is compared against the reference binary, and the reference
 * sampler can be timed on the same box.  The random stream itself stays "parity unpinned" (DESIGN.md section 6).
 * TEST INFRASTRUCTURE (oracle/_ref build); never part of libwholegraph.so.
 */
#pragma once
#include <cstdint>

#include <raft/random/rng_state.hpp>

#if defined(__CUDACC__)
#define WGREF_HDI __host__ __device__ inline
#else
#define WGREF_HDI inline
#endif

namespace raft {
namespace random {
namespace detail {

template <typename GenType>
struct DeviceState {
  using gen_t = GenType;
  explicit DeviceState(const RngState& rng_state) : seed(rng_state.seed), base_subsequence(rng_state.base_subsequence) {}
  uint64_t seed;
  uint64_t base_subsequence;
};

struct PCGenerator {
  WGREF_HDI PCGenerator(uint64_t seed, uint64_t subsequence, uint64_t /*offset, always 0 on this path*/) { init(seed, subsequence); }
  WGREF_HDI PCGenerator(const DeviceState<PCGenerator>& rng_state, const uint64_t subsequence)
  {
    init(rng_state.seed, rng_state.base_subsequence + subsequence);
  }

  WGREF_HDI void init(uint64_t seed, uint64_t subsequence)
  {
    pcg_state = 0;
    inc       = (subsequence << 1u) | 1u;
    next_u32();
    pcg_state += seed;
    next_u32();
  }

  WGREF_HDI uint32_t next_u32()
  {
    const uint64_t old  = pcg_state;
    pcg_state           = old * 6364136223846793005ULL + inc;
    const uint32_t x    = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot  = (uint32_t)(old >> 59u);
    return (x >> rot) | (x << ((0u - rot) & 31u));
  }
  WGREF_HDI uint64_t next_u64()
  {
    const uint64_t lo = next_u32();
    const uint64_t hi = next_u32();
    return lo | (hi << 32);
  }

  WGREF_HDI void next(uint32_t& ret) { ret = next_u32(); }
  WGREF_HDI void next(uint64_t& ret) { ret = next_u64(); }
  WGREF_HDI void next(int32_t& ret) { ret = (int32_t)(next_u32() & 0x7fffffffu); }
  WGREF_HDI void next(int64_t& ret) { ret = (int64_t)(next_u64() & 0x7fffffffffffffffULL); }
  WGREF_HDI void next(float& ret) { ret = (float)(next_u32() >> 8) / 16777216.0f; }
  WGREF_HDI void next(double& ret) { ret = (double)(next_u64() >> 11) / 9007199254740992.0; }

 private:
  uint64_t pcg_state;
  uint64_t inc;
};

template <typename Type>
struct UniformDistParams {
  Type start;
  Type end;
};

template <typename OutType, typename LenType = int, typename GenType>
WGREF_HDI void custom_next(GenType& gen, OutType* val, UniformDistParams<OutType> params, LenType /*idx*/ = 0, LenType /*stride*/ = 0)
{
  OutType res;
  gen.next(res);
  *val = (res * (params.end - params.start)) + params.start;
}

}  // namespace detail
}  // namespace random
}  // namespace raft
