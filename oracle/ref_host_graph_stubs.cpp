/* TEST INFRASTRUCTURE (oracle/_ref/ref_host_graph.so, CPU only).  The reference's sampling and graph-op entry points
 * (cpp/src/wholegraph_ops/{un,}weighted_sample_without_replacement.cpp, cpp/src/graph_ops/append_unique.cpp,
 * csr_add_self_loop.cpp) are compiled for the CPU as they are; the GPU functions they dispatch to are replaced here by
 * stubs returning the sentinel 1000 = "argument checks passed, call dispatched".  Signatures come from the reference's own
 * internal headers: a mismatch is a compile error. */
#include <cuda_runtime_api.h>

#include "graph_ops/append_unique_impl.h"
#include "graph_ops/csr_add_self_loop_impl.h"
#include "wholegraph_ops/unweighted_sample_without_replacement_impl.h"
#include "wholegraph_ops/weighted_sample_without_replacement_impl.h"

static const wholememory_error_code_t kDispatched = static_cast<wholememory_error_code_t>(1000);

namespace wholegraph_ops {
wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement_mapped(
  wholememory_gref_t, wholememory_array_description_t, wholememory_gref_t, wholememory_array_description_t, void*,
  wholememory_array_description_t, int, void*, wholememory_array_description_t, void*, void*, void*, unsigned long long,
  wholememory_env_func_t*, cudaStream_t)
{
  return kDispatched;
}
wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement_nccl(
  wholememory_handle_t, wholememory_handle_t, wholememory_tensor_description_t, wholememory_tensor_description_t, void*,
  wholememory_array_description_t, int, void*, wholememory_array_description_t, void*, void*, void*, unsigned long long,
  wholememory_env_func_t*, cudaStream_t)
{
  return kDispatched;
}
wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement_mapped(
  wholememory_gref_t, wholememory_array_description_t, wholememory_gref_t, wholememory_array_description_t, wholememory_gref_t,
  wholememory_array_description_t, void*, wholememory_array_description_t, int, void*, wholememory_array_description_t, void*, void*,
  void*, unsigned long long, wholememory_env_func_t*, cudaStream_t)
{
  return kDispatched;
}
}  // namespace wholegraph_ops

namespace graph_ops {
wholememory_error_code_t graph_append_unique_impl(void*, wholememory_array_description_t, void*, wholememory_array_description_t, void*, int*,
                                                  wholememory_env_func_t*, cudaStream_t)
{
  return kDispatched;
}
wholememory_error_code_t csr_add_self_loop_impl(int*, wholememory_array_description_t, int*, wholememory_array_description_t, int*,
                                                wholememory_array_description_t, int*, wholememory_array_description_t, cudaStream_t)
{
  return kDispatched;
}
}  // namespace graph_ops
