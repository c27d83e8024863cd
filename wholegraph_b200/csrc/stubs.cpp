/*
 * Entry points the reference's binding links against but which are outside this build's scope
 * (SURVEY section 8(f) "next": weighted sampling).  They exist so that the cython binding / ctypes loader resolves every symbol, and
 * they fail loudly with WHOLEMEMORY_NOT_IMPLEMENTED instead of silently doing nothing.
 */
#include "wm_internal.hpp"

extern "C" {

wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(wholememory_tensor_t,
                                                                            wholememory_tensor_t,
                                                                            wholememory_tensor_t,
                                                                            wholememory_tensor_t,
                                                                            int,
                                                                            wholememory_tensor_t,
                                                                            void*,
                                                                            void*,
                                                                            void*,
                                                                            unsigned long long,
                                                                            wholememory_env_func_t*,
                                                                            void*)
{
  WM_ERROR("weighted neighbor sampling is not built yet (SURVEY 8(f) rank 4)");
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}

wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(int64_t, int64_t, wholememory_tensor_t)
{
  WM_ERROR("generate_exponential_distribution_negative_float_cpu belongs to weighted sampling (not built yet)");
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}

} /* extern "C" */
