/* TEST INFRASTRUCTURE (oracle/_ref/ref_host_sampling_model.so, CPU only).  extern "C" doors to the reference's OWN CPU
 * models of the neighbor samplers -- the functions its GPU tests compare the kernels against:
 *   wholegraph_ops::testing::wholegraph_csr_unweighted_sample_without_replacement_cpu
 *   wholegraph_ops::testing::wholegraph_csr_weighted_sample_without_replacement_cpu
 * (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:419-505, :676-763), compiled from the reference tree as they are
 * on top of the restated PCG stand-in.  tests/test_ref_sampling_model.py checks this repo's oracle against them. */
#include <cstdlib>
#include <cstring>

#include "wholegraph_ops/graph_sampling_test_utils.hpp"

extern "C" {

/* outputs are malloc'ed by the reference code; release them with wgref_model_free */
int wgref_cpu_unweighted_sample(void* row_ptr, int64_t row_ptr_count, void* col, int col_is_int64, int64_t col_count, void* centers,
                                int centers_is_int64, int64_t center_count, int max_sample_count, unsigned long long seed, void** offsets,
                                void** dst, void** center_lid, void** edge_gid)
{
  auto rp = wholememory_create_array_desc(row_ptr_count, 0, WHOLEMEMORY_DT_INT64);
  auto cp = wholememory_create_array_desc(col_count, 0, col_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  auto cn = wholememory_create_array_desc(center_count, 0, centers_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  auto od = wholememory_create_array_desc(center_count + 1, 0, WHOLEMEMORY_DT_INT);
  int total = 0;
  wholegraph_ops::testing::wholegraph_csr_unweighted_sample_without_replacement_cpu(row_ptr, rp, col, cp, centers, cn, max_sample_count, offsets, od,
                                                                                    dst, center_lid, edge_gid, &total, seed);
  return total;
}

int wgref_cpu_weighted_sample(void* row_ptr, int64_t row_ptr_count, void* col, int col_is_int64, int64_t col_count, void* weights,
                              int weights_is_double, void* centers, int centers_is_int64, int64_t center_count, int max_sample_count,
                              unsigned long long seed, void** offsets, void** dst, void** center_lid, void** edge_gid)
{
  auto rp = wholememory_create_array_desc(row_ptr_count, 0, WHOLEMEMORY_DT_INT64);
  auto cp = wholememory_create_array_desc(col_count, 0, col_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  auto wp = wholememory_create_array_desc(col_count, 0, weights_is_double ? WHOLEMEMORY_DT_DOUBLE : WHOLEMEMORY_DT_FLOAT);
  auto cn = wholememory_create_array_desc(center_count, 0, centers_is_int64 ? WHOLEMEMORY_DT_INT64 : WHOLEMEMORY_DT_INT);
  auto od = wholememory_create_array_desc(center_count + 1, 0, WHOLEMEMORY_DT_INT);
  int total = 0;
  wholegraph_ops::testing::wholegraph_csr_weighted_sample_without_replacement_cpu(row_ptr, rp, col, cp, weights, wp, centers, cn, max_sample_count,
                                                                                  offsets, od, dst, center_lid, edge_gid, &total, seed);
  return total;
}

void wgref_model_free(void* p) { free(p); }

} /* extern "C" */
