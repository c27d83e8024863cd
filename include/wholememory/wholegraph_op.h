/*
 * CSR neighbor sampling out of WholeMemory.
 * Drop-in for reference cpp/include/wholememory/wholegraph_op.h:43-105 (impl there:
 * cpp/src/wholegraph_ops/unweighted_sample_without_replacement.cpp:23 and _func.cuh:284-475).
 * Implementation here: wholegraph_b200/csrc/neighbor_sample.cu.
 */
#pragma once
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * For every center node c: take min(degree(c), max_sample_count) distinct neighbors
 * (all of them, in CSR order, when degree <= max_sample_count or max_sample_count <= 0).
 *   wm_csr_row_ptr_tensor  : int64 [num_nodes+1]    WholeMemory (CONTINUOUS/CHUNKED) or device ptr
 *   wm_csr_col_ptr_tensor  : int32|int64 [num_edges]
 *   center_nodes_tensor    : int32|int64 [n] device array
 *   output_sample_offset_tensor : int32 [n+1] device array (exclusive scan of sample counts)
 *   output_*_memory_context: caller contexts filled through env_fns->output_fns.malloc_fn with
 *       dst node ids (col dtype), center local ids (int32), edge global ids (int64); the last two
 *       may be nullptr.
 */
wholememory_error_code_t wholegraph_csr_unweighted_sample_without_replacement(
    wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
    wholememory_tensor_t center_nodes_tensor, int max_sample_count, wholememory_tensor_t output_sample_offset_tensor,
    void* output_dest_memory_context, void* output_center_localid_memory_context,
    void* output_edge_gid_memory_context, unsigned long long random_seed, wholememory_env_func_t* env_fns,
    void* stream);

/* weighted variant (float/double edge weights): SURVEY section 8(f) rank 4 */
wholememory_error_code_t wholegraph_csr_weighted_sample_without_replacement(
    wholememory_tensor_t wm_csr_row_ptr_tensor, wholememory_tensor_t wm_csr_col_ptr_tensor,
    wholememory_tensor_t wm_csr_weight_ptr_tensor, wholememory_tensor_t center_nodes_tensor, int max_sample_count,
    wholememory_tensor_t output_sample_offset_tensor, void* output_dest_memory_context,
    void* output_center_localid_memory_context, void* output_edge_gid_memory_context, unsigned long long random_seed,
    wholememory_env_func_t* env_fns, void* stream);

/* host replay of the sampler's random stream, used by the tests to rebuild expected samples:
 * output[i] = i-th positive int32 / negative-exponential float drawn by (seed, subsequence). */
wholememory_error_code_t generate_random_positive_int_cpu(int64_t random_seed, int64_t subsequence,
    wholememory_tensor_t output);
wholememory_error_code_t generate_exponential_distribution_negative_float_cpu(int64_t random_seed,
    int64_t subsequence, wholememory_tensor_t output);

#ifdef __cplusplus
}
#endif
