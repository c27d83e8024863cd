/*
 * Stand-in for rapidsai/raft branch-24.12 raft/random/rng_state.hpp (RAFT is not vendored here): the three-argument
 * RngState the reference constructs (cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:368,
 * raft_random_gen.cu:43).  TEST INFRASTRUCTURE (oracle/_ref build); restated, not copied -- see rng_device.cuh.
 */
#pragma once
#include <cstdint>

namespace raft {
namespace random {

enum GeneratorType { GenPhilox = 0, GenPC };

struct RngState {
  explicit RngState(uint64_t _seed) : seed(_seed) {}
  RngState(uint64_t _seed, GeneratorType _type) : seed(_seed), type(_type) {}
  RngState(uint64_t _seed, uint64_t _base_subsequence, GeneratorType _type) : seed(_seed), base_subsequence(_base_subsequence), type(_type) {}
  uint64_t seed{0};
  uint64_t base_subsequence{0};
  GeneratorType type{GenPhilox};
};

}  // namespace random
}  // namespace raft
