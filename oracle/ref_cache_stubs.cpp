/* TEST INFRASTRUCTURE (oracle/_ref build only).  The reference's embedding layer (cpp/src/wholememory/embedding.cpp,
 * embedding_optimizer.cpp, embedding_cache.cpp) is compiled from /root/reference as it is; the five device-cache entry
 * points it calls live in TUs that need RAFT's real select_k (embedding_cache_func.cu, gather_cached_func.cu) and cannot
 * be built here.  They are only reached when an embedding is created WITH a cache policy, which nothing in this repo's
 * tests or benches does, so they are stubbed to fail loudly.  Signatures come from the reference's own headers
 * (functions/embedding_cache_func.h:40-81, functions/gather_cached_func.h:26-52): a mismatch is a compile error. */
#include <cstdio>

#include "wholememory_ops/functions/embedding_cache_func.h"
#include "wholememory_ops/functions/gather_cached_func.h"

namespace wholememory_ops {

static wholememory_error_code_t refuse(const char* what)
{
  fprintf(stderr, "[oracle/_ref] %s is not built (needs RAFT select_k); cached embeddings are unavailable in this reference build\n", what);
  return WHOLEMEMORY_NOT_IMPLEMENTED;
}

wholememory_error_code_t update_cache_direct_same_comm(void*, wholememory_array_description_t, wholememory_tensor_t,
                                                       const wholememory::embedding_cache_local_data*, int, wholememory_env_func_t*,
                                                       cudaStream_t)
{
  return refuse("update_cache_direct_same_comm");
}

wholememory_error_code_t update_cache_different_comm(void*, wholememory_array_description_t, wholememory_tensor_t, wholememory_comm_t,
                                                     size_t*, const wholememory::embedding_cache_local_data*, int,
                                                     wholememory_env_func_t*, cudaStream_t)
{
  return refuse("update_cache_different_comm");
}

wholememory_error_code_t writeback_cache_direct_same_comm(wholememory_tensor_t, const wholememory::embedding_cache_local_data*, int, bool,
                                                          cudaStream_t)
{
  return refuse("writeback_cache_direct_same_comm");
}

wholememory_error_code_t gather_cached_func(wholememory_gref_t, wholememory_tensor_description_t*, wholememory_gref_t,
                                            wholememory_tensor_description_t*, wholememory_gref_t, void*, wholememory_tensor_description_t*,
                                            void*, wholememory_tensor_description_t*, int, int64_t, int64_t, cudaStream_t)
{
  return refuse("gather_cached_func");
}

wholememory_error_code_t try_gather_cached_func(wholememory_gref_t, wholememory_tensor_description_t*, wholememory_gref_t, void*,
                                                wholememory_tensor_description_t*, void*, void*, void*, wholememory_tensor_description_t*,
                                                int, int64_t, cudaStream_t)
{
  return refuse("try_gather_cached_func");
}

}  // namespace wholememory_ops
