/*
 * Tensor / matrix / array descriptors of the WholeMemory C ABI.
 * Enum values and struct layouts are ABI: they match reference
 * cpp/include/wholememory/tensor_description.h:29-99 (dtype enum :29-40, array :60-64,
 * matrix :69-74, tensor :81-99 -- 160 bytes).  Helper semantics follow
 * cpp/src/wholememory/tensor_description.cpp:20-233.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wholememory_dtype_t {
  WHOLEMEMORY_DT_UNKNOWN = 0,
  WHOLEMEMORY_DT_FLOAT,  /* fp32 */
  WHOLEMEMORY_DT_HALF,   /* fp16 */
  WHOLEMEMORY_DT_DOUBLE, /* fp64 */
  WHOLEMEMORY_DT_BF16,
  WHOLEMEMORY_DT_INT, /* int32 */
  WHOLEMEMORY_DT_INT64,
  WHOLEMEMORY_DT_INT16,
  WHOLEMEMORY_DT_INT8,
  WHOLEMEMORY_DT_COUNT,
};

size_t wholememory_dtype_get_element_size(wholememory_dtype_t dtype);
bool wholememory_dtype_is_floating_number(wholememory_dtype_t dtype);
bool wholememory_dtype_is_integer_number(wholememory_dtype_t dtype);

/* all sizes / strides / offsets below are in ELEMENTS, never bytes */
struct wholememory_array_description_t {
  int64_t size;
  int64_t storage_offset;
  wholememory_dtype_t dtype;
};

struct wholememory_matrix_description_t {
  int64_t sizes[2]; /* rows, columns */
  int64_t stride;   /* row stride */
  int64_t storage_offset;
  wholememory_dtype_t dtype;
};

#define WHOLEMEMORY_MAX_TENSOR_DIM (8)

struct wholememory_tensor_description_t {
  int64_t sizes[WHOLEMEMORY_MAX_TENSOR_DIM];
  int64_t strides[WHOLEMEMORY_MAX_TENSOR_DIM];
  int64_t storage_offset;
  int dim;
  wholememory_dtype_t dtype;
};

wholememory_array_description_t wholememory_create_array_desc(int64_t size,
                                                              int64_t storage_offset,
                                                              wholememory_dtype_t dtype);
wholememory_matrix_description_t wholememory_create_matrix_desc(int64_t sizes[2],
                                                                int64_t stride,
                                                                int64_t storage_offset,
                                                                wholememory_dtype_t dtype);
/* dim=0, all sizes/strides 1, offset 0, dtype UNKNOWN */
void wholememory_initialize_tensor_desc(wholememory_tensor_description_t* desc);

void wholememory_copy_array_desc_to_matrix(wholememory_matrix_description_t* m,
                                           wholememory_array_description_t* a);
void wholememory_copy_array_desc_to_tensor(wholememory_tensor_description_t* t,
                                           wholememory_array_description_t* a);
void wholememory_copy_matrix_desc_to_tensor(wholememory_tensor_description_t* t,
                                            wholememory_matrix_description_t* m);
/* return false when the tensor is not expressible as the requested view */
bool wholememory_convert_tensor_desc_to_array(wholememory_array_description_t* a,
                                              wholememory_tensor_description_t* t);
bool wholememory_convert_tensor_desc_to_matrix(wholememory_matrix_description_t* m,
                                               wholememory_tensor_description_t* t);

int64_t wholememory_get_memory_element_count_from_array(wholememory_array_description_t* a);
int64_t wholememory_get_memory_size_from_array(wholememory_array_description_t* a);
int64_t wholememory_get_memory_element_count_from_matrix(wholememory_matrix_description_t* m);
int64_t wholememory_get_memory_size_from_matrix(wholememory_matrix_description_t* m);
int64_t wholememory_get_memory_element_count_from_tensor(wholememory_tensor_description_t* t);
int64_t wholememory_get_memory_size_from_tensor(wholememory_tensor_description_t* t);

bool wholememory_squeeze_tensor(wholememory_tensor_description_t* t, int dim);
bool wholememory_unsqueeze_tensor(wholememory_tensor_description_t* t, int dim);

#ifdef __cplusplus
}
#endif
