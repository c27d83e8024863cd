/* TEST INFRASTRUCTURE (oracle/_ref/ref_host_ops.so, CPU only).  The reference's op entry points
 * (cpp/src/wholememory_ops/gather_op.cpp, scatter_op.cpp) are compiled for the CPU as they are; the functions they
 * dispatch to live in GPU translation units, so they are replaced here by stubs that report "the argument checks passed
 * and the call was dispatched" through a sentinel code no real path returns.  Signatures come from the reference's own
 * internal headers (gather_op_impl.h, scatter_op_impl.h): a mismatch is a compile error. */
#include <cuda_runtime_api.h>
#include <wholememory/env_func_ptrs.h>

#include "wholememory_ops/gather_op_impl.h"
#include "wholememory_ops/scatter_op_impl.h"

static const wholememory_error_code_t kDispatched = static_cast<wholememory_error_code_t>(1000);

namespace wholememory_ops {

wholememory_error_code_t wholememory_gather_mapped(wholememory_gref_t, wholememory_matrix_description_t, void*, wholememory_array_description_t,
                                                   void*, wholememory_matrix_description_t, bool, wholememory_env_func_t*, cudaStream_t, int)
{
  return kDispatched;
}
wholememory_error_code_t wholememory_gather_distributed(wholememory_handle_t, wholememory_matrix_description_t, void*,
                                                        wholememory_array_description_t, void*, wholememory_matrix_description_t,
                                                        wholememory_env_func_t*, cudaStream_t, int)
{
  return kDispatched;
}
wholememory_error_code_t wholememory_gather_hierarchy(wholememory_handle_t, wholememory_matrix_description_t, void*,
                                                      wholememory_array_description_t, void*, wholememory_matrix_description_t,
                                                      wholememory_env_func_t*, cudaStream_t, int)
{
  return kDispatched;
}
wholememory_error_code_t wholememory_scatter_mapped(void*, wholememory_matrix_description_t, void*, wholememory_array_description_t,
                                                    wholememory_gref_t, wholememory_matrix_description_t, wholememory_env_func_t*, cudaStream_t,
                                                    int)
{
  return kDispatched;
}
wholememory_error_code_t wholememory_scatter_distributed(void*, wholememory_matrix_description_t, void*, wholememory_array_description_t,
                                                         wholememory_handle_t, wholememory_matrix_description_t, wholememory_env_func_t*,
                                                         cudaStream_t, int)
{
  return kDispatched;
}

}  // namespace wholememory_ops
