/*
 * WholeMemory tensor views (1-D / 2-D strided views over a WholeMemory handle or a raw pointer).
 * Drop-in for reference cpp/include/wholememory/wholememory_tensor.h:40-190; behaviour follows
 * cpp/src/wholememory/wm_tensor.cpp:51-469.  Implementation: wholegraph_b200/csrc/tensor.cpp.
 */
#pragma once
#include <wholememory/tensor_description.h>
#include <wholememory/wholememory.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wholememory_tensor_* wholememory_tensor_t;

/* Collective: allocates sizes[0]*strides[0] elements; partition granularity = one row.
 * desc must be 1-D/2-D, storage_offset 0, innermost stride 1. */
wholememory_error_code_t wholememory_create_tensor(wholememory_tensor_t* wm_tensor,
                                                   wholememory_tensor_description_t* tensor_description,
                                                   wholememory_comm_t comm,
                                                   wholememory_memory_type_t memory_type,
                                                   wholememory_memory_location_t memory_location,
                                                   size_t* tensor_entry_partition = nullptr);
/* frees the handle too when the tensor owns it (collective in that case) */
wholememory_error_code_t wholememory_destroy_tensor(wholememory_tensor_t wm_tensor);
/* non-owning view of caller memory (indices, outputs, gradients...) */
wholememory_error_code_t wholememory_make_tensor_from_pointer(
  wholememory_tensor_t* wm_tensor,
  void* storage_ptr,
  wholememory_tensor_description_t* tensor_description);
/* non-owning view of an existing handle */
wholememory_error_code_t wholememory_make_tensor_from_handle(
  wholememory_tensor_t* wm_tensor,
  wholememory_handle_t handle,
  wholememory_tensor_description_t* tensor_description);

bool wholememory_tensor_has_handle(wholememory_tensor_t wm_tensor);
wholememory_handle_t wholememory_tensor_get_memory_handle(wholememory_tensor_t wm_tensor);
wholememory_tensor_description_t* wholememory_tensor_get_tensor_description(
  wholememory_tensor_t wm_tensor);
wholememory_error_code_t wholememory_tensor_get_global_reference(
  wholememory_tensor_t wm_tensor, wholememory_gref_t* gref);
/* view of this rank's rows as a plain pointer tensor */
wholememory_error_code_t wholememory_tensor_map_local_tensor(wholememory_tensor_t wm_tensor,
                                                             wholememory_tensor_t* local_tensor);
/* first element address (storage_offset applied); NULL for handle tensors that are not CONTINUOUS */
void* wholememory_tensor_get_data_pointer(wholememory_tensor_t wm_tensor);
/* world_size+1 row offsets / world_size row counts of the partition */
wholememory_error_code_t wholememory_tensor_get_entry_offsets(size_t* entry_offsets,
                                                              wholememory_tensor_t wm_tensor);
wholememory_error_code_t wholememory_tensor_get_entry_partition_sizes(
  size_t* entry_partition, wholememory_tensor_t wm_tensor);
wholememory_error_code_t wholememory_tensor_get_local_entry_count(
  size_t* local_entry_count, wholememory_tensor_t wm_tensor);
wholememory_error_code_t wholememory_tensor_get_local_entry_start(
  size_t* local_entry_start, wholememory_tensor_t wm_tensor);
/* [starts, ends) per dim, -1 = from begin / to end; non-owning */
wholememory_error_code_t wholememory_tensor_get_subtensor(wholememory_tensor_t wm_tensor,
                                                          int64_t* starts,
                                                          int64_t* ends,
                                                          wholememory_tensor_t* sub_wholememory_tensor);
wholememory_tensor_t wholememory_tensor_get_root(wholememory_tensor_t wm_tensor);

#define WM_TENSOR_COUNT_DEBUG
/* live tensor objects (leak check used by the Python tests) */
int64_t get_wholememory_tensor_count();

#ifdef __cplusplus
}
#endif
