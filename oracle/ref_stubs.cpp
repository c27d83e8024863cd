/* Stubs for the parts of the reference library that cannot be built here (they include un-vendored RAFT
 * headers or are out of the hot path): hierarchy gather, file I/O.  TEST INFRASTRUCTURE (oracle/_ref build).
 * Signatures follow the reference's internal headers (gather_op_impl.h, file_io.h). */
#include <cuda_runtime_api.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/tensor_description.h>

namespace wholememory_ops {
wholememory_error_code_t wholememory_gather_hierarchy(wholememory_handle_t,
                                                      wholememory_matrix_description_t,
                                                      void*,
                                                      wholememory_array_description_t,
                                                      void*,
                                                      wholememory_matrix_description_t,
                                                      wholememory_env_func_t*,
                                                      cudaStream_t,
                                                      int)
{
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}
}  // namespace wholememory_ops

namespace wholememory {
wholememory_error_code_t load_file_to_handle(wholememory_handle_t, size_t, size_t, size_t, const char**, int, int) noexcept
{
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}
wholememory_error_code_t store_handle_to_file(wholememory_handle_t, size_t, size_t, size_t, const char*) noexcept
{
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}
}  // namespace wholememory

/* cpp/src/parallel_utils.cpp:342 (fork harness, not built): device count for fork_get_device_count() */
int ForkGetDeviceCount()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
