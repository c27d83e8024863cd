/*
 * Graph helper ops that sit between sampling hops (reference cpp/include/wholememory/graph_op.h:39-59).
 * Implementation: wholegraph_b200/csrc/graph_ops.cu (first-occurrence-ordered append-unique, CSR self-loop insert).
 */
#pragma once
#include <cuda_runtime_api.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/* unique(targets ++ neighbors) with targets first, plus neighbor -> unique-position map */
wholememory_error_code_t graph_append_unique(wholememory_tensor_t target_nodes_tensor,
    wholememory_tensor_t neighbor_nodes_tensor, void* output_unique_node_memory_context,
    wholememory_tensor_t output_neighbor_raw_to_unique_mapping_tensor, wholememory_env_func_t* env_fns, void* stream);

/* CSR + one self edge per row (placed first) */
wholememory_error_code_t csr_add_self_loop(wholememory_tensor_t csr_row_ptr_tensor,
    wholememory_tensor_t csr_col_ptr_tensor, wholememory_tensor_t output_csr_row_ptr_tensor,
    wholememory_tensor_t output_csr_col_ptr_tensor, void* stream);

#ifdef __cplusplus
}
#endif
