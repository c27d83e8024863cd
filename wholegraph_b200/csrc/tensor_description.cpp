/*
 * Descriptor helpers of the C ABI (pure host functions).
 * Behaviour follows reference cpp/src/wholememory/tensor_description.cpp:20-233.
 */
#include <wholememory/global_reference.h>
#include <wholememory/tensor_description.h>

namespace {
inline bool known_dtype(wholememory_dtype_t d) { return d > WHOLEMEMORY_DT_UNKNOWN && d < WHOLEMEMORY_DT_COUNT; }
}  // namespace

extern "C" {

size_t wholememory_dtype_get_element_size(wholememory_dtype_t dtype)
{
  static const int bytes[WHOLEMEMORY_DT_COUNT] = {
    /*UNKNOWN*/ 0, /*FLOAT*/ 4, /*HALF*/ 2, /*DOUBLE*/ 8, /*BF16*/ 2,
    /*INT*/ 4,     /*INT64*/ 8, /*INT16*/ 2, /*INT8*/ 1};
  if (dtype < 0 || dtype >= WHOLEMEMORY_DT_COUNT) return static_cast<size_t>(-1);
  return bytes[dtype];
}

bool wholememory_dtype_is_floating_number(wholememory_dtype_t dtype)
{
  switch (dtype) {
    case WHOLEMEMORY_DT_FLOAT:
    case WHOLEMEMORY_DT_HALF:
    case WHOLEMEMORY_DT_DOUBLE:
    case WHOLEMEMORY_DT_BF16: return true;
    default: return false;
  }
}

bool wholememory_dtype_is_integer_number(wholememory_dtype_t dtype)
{
  switch (dtype) {
    case WHOLEMEMORY_DT_INT:
    case WHOLEMEMORY_DT_INT64:
    case WHOLEMEMORY_DT_INT16:
    case WHOLEMEMORY_DT_INT8: return true;
    default: return false;
  }
}

wholememory_array_description_t wholememory_create_array_desc(int64_t size,
                                                              int64_t storage_offset,
                                                              wholememory_dtype_t dtype)
{
  return wholememory_array_description_t{size, storage_offset, dtype};
}

wholememory_matrix_description_t wholememory_create_matrix_desc(int64_t sizes[2],
                                                                int64_t stride,
                                                                int64_t storage_offset,
                                                                wholememory_dtype_t dtype)
{
  return wholememory_matrix_description_t{{sizes[0], sizes[1]}, stride, storage_offset, dtype};
}

void wholememory_initialize_tensor_desc(wholememory_tensor_description_t* d)
{
  for (int i = 0; i < WHOLEMEMORY_MAX_TENSOR_DIM; ++i) d->sizes[i] = d->strides[i] = 1;
  d->storage_offset = 0;
  d->dim            = 0;
  d->dtype          = WHOLEMEMORY_DT_UNKNOWN;
}

void wholememory_copy_array_desc_to_matrix(wholememory_matrix_description_t* m,
                                           wholememory_array_description_t* a)
{
  *m = wholememory_matrix_description_t{{a->size, 1}, 1, a->storage_offset, a->dtype};
}

void wholememory_copy_array_desc_to_tensor(wholememory_tensor_description_t* t,
                                           wholememory_array_description_t* a)
{
  wholememory_initialize_tensor_desc(t);
  t->dim            = 1;
  t->sizes[0]       = a->size;
  t->storage_offset = a->storage_offset;
  t->dtype          = a->dtype;
}

void wholememory_copy_matrix_desc_to_tensor(wholememory_tensor_description_t* t,
                                            wholememory_matrix_description_t* m)
{
  wholememory_initialize_tensor_desc(t);
  t->dim            = 2;
  t->sizes[0]       = m->sizes[0];
  t->sizes[1]       = m->sizes[1];
  t->strides[0]     = m->stride;
  t->storage_offset = m->storage_offset;
  t->dtype          = m->dtype;
}

bool wholememory_convert_tensor_desc_to_array(wholememory_array_description_t* a,
                                              wholememory_tensor_description_t* t)
{
  if (!known_dtype(t->dtype) || t->dim != 1 || t->strides[0] != 1) return false;
  *a = wholememory_array_description_t{t->sizes[0], t->storage_offset, t->dtype};
  return true;
}

bool wholememory_convert_tensor_desc_to_matrix(wholememory_matrix_description_t* m,
                                               wholememory_tensor_description_t* t)
{
  if (!known_dtype(t->dtype) || t->dim < 1 || t->dim > 2) return false;
  if (t->dim == 2 && t->strides[1] != 1) return false;
  m->dtype          = t->dtype;
  m->storage_offset = t->storage_offset;
  m->sizes[0]       = t->sizes[0];
  m->sizes[1]       = t->dim == 2 ? t->sizes[1] : 1;
  m->stride         = t->dim == 2 ? t->strides[0] : 1;
  return true;
}

int64_t wholememory_get_memory_element_count_from_array(wholememory_array_description_t* a) { return a->size; }

int64_t wholememory_get_memory_size_from_array(wholememory_array_description_t* a)
{
  return a->size * (int64_t)wholememory_dtype_get_element_size(a->dtype);
}

int64_t wholememory_get_memory_element_count_from_matrix(wholememory_matrix_description_t* m)
{
  return m->sizes[0] * m->stride;
}

int64_t wholememory_get_memory_size_from_matrix(wholememory_matrix_description_t* m)
{
  return m->sizes[0] * m->stride * (int64_t)wholememory_dtype_get_element_size(m->dtype);
}

int64_t wholememory_get_memory_element_count_from_tensor(wholememory_tensor_description_t* t)
{
  if (t->dim == 0) return 1;
  if (t->dim < 0 || t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return -1;
  return t->sizes[0] * t->strides[0];
}

int64_t wholememory_get_memory_size_from_tensor(wholememory_tensor_description_t* t)
{
  return wholememory_get_memory_element_count_from_tensor(t) *
         (int64_t)wholememory_dtype_get_element_size(t->dtype);
}

bool wholememory_squeeze_tensor(wholememory_tensor_description_t* t, int dim)
{
  if (t == nullptr || dim < 0 || dim >= t->dim || t->sizes[dim] != 1) return false;
  /* an inner unit dim may only go when it does not carry a distinct stride */
  if (dim != t->dim - 1 && t->strides[dim] != t->strides[dim + 1]) return false;
  for (int i = dim; i + 1 < t->dim; ++i) {
    t->sizes[i]   = t->sizes[i + 1];
    t->strides[i] = t->strides[i + 1];
  }
  --t->dim;
  return true;
}

bool wholememory_unsqueeze_tensor(wholememory_tensor_description_t* t, int dim)
{
  if (t == nullptr || dim < 0 || dim > t->dim || t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return false;
  /* the new unit dim inherits the stride of the dim it is inserted before (or the last stride) */
  int64_t inherited = t->dim > 0 ? t->strides[t->dim - 1] : 1;
  for (int i = t->dim; i > dim; --i) {
    t->sizes[i]   = t->sizes[i - 1];
    t->strides[i] = t->strides[i - 1];
    inherited     = t->strides[i];
  }
  t->sizes[dim]   = 1;
  t->strides[dim] = inherited;
  ++t->dim;
  return true;
}

wholememory_gref_t wholememory_create_continuous_global_reference(void* ptr)
{
  wholememory_gref_t g;
  g.pointer             = ptr;
  g.rank_memory_offsets = nullptr;
  g.world_size          = 1;
  g.stride              = 0;
  g.same_chunk          = true;
  return g;
}

} /* extern "C" */
