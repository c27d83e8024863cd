/*
 * Array / matrix / tensor descriptors of the C ABI: pure host arithmetic, no device work.
 * The behaviour is the reference's (cpp/src/wholememory/tensor_description.cpp:20-233); tests/test_ref_host_tensor.py
 * runs 200,000 random cases through that file compiled for the CPU and through this one and finds no divergence.
 *
 * Layout of this file: one trait table for the dtypes, one helper that zeroes a tensor descriptor, then the exported
 * functions grouped as constructors, conversions, footprints and the two reshapes.
 */
#include <wholememory/global_reference.h>
#include <wholememory/tensor_description.h>

namespace {

using adesc = wholememory_array_description_t;
using mdesc = wholememory_matrix_description_t;
using tdesc = wholememory_tensor_description_t;
using dt    = wholememory_dtype_t;

struct dtype_traits {
  int bytes;
  bool floating;
  bool integer;
};
/* indexed by wholememory_dtype_t: UNKNOWN, FLOAT, HALF, DOUBLE, BF16, INT, INT64, INT16, INT8 */
constexpr dtype_traits kTraits[WHOLEMEMORY_DT_COUNT] = {{0, false, false},
                                                        {4, true, false},
                                                        {2, true, false},
                                                        {8, true, false},
                                                        {2, true, false},
                                                        {4, false, true},
                                                        {8, false, true},
                                                        {2, false, true},
                                                        {1, false, true}};

inline bool in_table(dt d) { return d >= 0 && d < WHOLEMEMORY_DT_COUNT; }
inline bool is_concrete(dt d) { return d > WHOLEMEMORY_DT_UNKNOWN && d < WHOLEMEMORY_DT_COUNT; }
inline int64_t width(dt d) { return static_cast<int64_t>(wholememory_dtype_get_element_size(d)); }

void blank(tdesc* t)
{
  t->dim            = 0;
  t->dtype          = WHOLEMEMORY_DT_UNKNOWN;
  t->storage_offset = 0;
  for (int axis = 0; axis < WHOLEMEMORY_MAX_TENSOR_DIM; ++axis) {
    t->sizes[axis]   = 1;
    t->strides[axis] = 1;
  }
}

/* move axes [from, dim) one slot towards the front, dropping axis from-1 ... used by squeeze */
void close_gap(tdesc* t, int at)
{
  for (int axis = at + 1; axis < t->dim; ++axis) {
    t->sizes[axis - 1]   = t->sizes[axis];
    t->strides[axis - 1] = t->strides[axis];
  }
  t->dim -= 1;
}

}  // namespace

extern "C" {

/* ---- dtypes ---- */
size_t wholememory_dtype_get_element_size(dt d) { return in_table(d) ? kTraits[d].bytes : static_cast<size_t>(-1); }
bool wholememory_dtype_is_floating_number(dt d) { return in_table(d) && kTraits[d].floating; }
bool wholememory_dtype_is_integer_number(dt d) { return in_table(d) && kTraits[d].integer; }

/* ---- constructors ---- */
adesc wholememory_create_array_desc(int64_t size, int64_t storage_offset, dt dtype)
{
  adesc a;
  a.dtype          = dtype;
  a.storage_offset = storage_offset;
  a.size           = size;
  return a;
}

mdesc wholememory_create_matrix_desc(int64_t sizes[2], int64_t stride, int64_t storage_offset, dt dtype)
{
  mdesc m;
  m.dtype          = dtype;
  m.storage_offset = storage_offset;
  m.stride         = stride;
  m.sizes[0]       = sizes[0];
  m.sizes[1]       = sizes[1];
  return m;
}

void wholememory_initialize_tensor_desc(tdesc* t) { blank(t); }

/* ---- conversions between the three descriptor kinds ---- */
void wholememory_copy_array_desc_to_matrix(mdesc* m, adesc* a)
{
  /* an array is a one-column matrix whose rows are adjacent */
  int64_t shape[2] = {a->size, 1};
  *m               = wholememory_create_matrix_desc(shape, 1, a->storage_offset, a->dtype);
}

void wholememory_copy_array_desc_to_tensor(tdesc* t, adesc* a)
{
  blank(t);
  t->dtype          = a->dtype;
  t->storage_offset = a->storage_offset;
  t->sizes[0]       = a->size;
  t->dim            = 1;
}

void wholememory_copy_matrix_desc_to_tensor(tdesc* t, mdesc* m)
{
  blank(t);
  t->dtype          = m->dtype;
  t->storage_offset = m->storage_offset;
  t->strides[0]     = m->stride;
  for (int axis = 0; axis < 2; ++axis) t->sizes[axis] = m->sizes[axis];
  t->dim = 2;
}

bool wholememory_convert_tensor_desc_to_array(adesc* a, tdesc* t)
{
  const bool dense_vector = t->dim == 1 && t->strides[0] == 1;
  if (!is_concrete(t->dtype) || !dense_vector) return false;
  *a = wholememory_create_array_desc(t->sizes[0], t->storage_offset, t->dtype);
  return true;
}

bool wholememory_convert_tensor_desc_to_matrix(mdesc* m, tdesc* t)
{
  if (!is_concrete(t->dtype)) return false;
  int64_t shape[2];
  int64_t row_pitch;
  if (t->dim == 1) { /* a vector reads as n x 1 */
    shape[0] = t->sizes[0], shape[1] = 1, row_pitch = 1;
  } else if (t->dim == 2 && t->strides[1] == 1) {
    shape[0] = t->sizes[0], shape[1] = t->sizes[1], row_pitch = t->strides[0];
  } else {
    return false;
  }
  *m = wholememory_create_matrix_desc(shape, row_pitch, t->storage_offset, t->dtype);
  return true;
}

/* ---- footprints: elements / bytes spanned from the first to the last addressed element's row end ---- */
int64_t wholememory_get_memory_element_count_from_array(adesc* a) { return a->size; }
int64_t wholememory_get_memory_size_from_array(adesc* a) { return width(a->dtype) * a->size; }

int64_t wholememory_get_memory_element_count_from_matrix(mdesc* m) { return m->stride * m->sizes[0]; }
int64_t wholememory_get_memory_size_from_matrix(mdesc* m)
{
  return width(m->dtype) * wholememory_get_memory_element_count_from_matrix(m);
}

int64_t wholememory_get_memory_element_count_from_tensor(tdesc* t)
{
  if (t->dim < 0 || t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return -1;
  return t->dim == 0 ? 1 : t->strides[0] * t->sizes[0];
}
int64_t wholememory_get_memory_size_from_tensor(tdesc* t)
{
  return width(t->dtype) * wholememory_get_memory_element_count_from_tensor(t);
}

/* ---- reshapes ---- */
bool wholememory_squeeze_tensor(tdesc* t, int dim)
{
  if (t == nullptr) return false;
  const bool axis_exists = dim >= 0 && dim < t->dim;
  if (!axis_exists || t->sizes[dim] != 1) return false;
  /* an inner unit axis may only go when it does not carry a stride of its own */
  const bool innermost = dim + 1 == t->dim;
  if (!innermost && t->strides[dim] != t->strides[dim + 1]) return false;
  close_gap(t, dim);
  return true;
}

bool wholememory_unsqueeze_tensor(tdesc* t, int dim)
{
  if (t == nullptr) return false;
  if (dim < 0 || dim > t->dim || t->dim >= WHOLEMEMORY_MAX_TENSOR_DIM) return false;
  /* open a slot at `dim`; the unit axis takes the stride of the axis that now follows it, or, when appended at the
   * end, the stride of the previously last axis (1 for a scalar) */
  int64_t pitch = t->dim > 0 ? t->strides[t->dim - 1] : 1;
  if (dim < t->dim) pitch = t->strides[dim];
  for (int axis = t->dim; axis > dim; --axis) {
    t->sizes[axis]   = t->sizes[axis - 1];
    t->strides[axis] = t->strides[axis - 1];
  }
  t->sizes[dim]   = 1;
  t->strides[dim] = pitch;
  t->dim += 1;
  return true;
}

/* ---- a plain pointer seen as a one-rank global reference ---- */
wholememory_gref_t wholememory_create_continuous_global_reference(void* ptr)
{
  wholememory_gref_t whole{};
  whole.pointer             = ptr;
  whole.rank_memory_offsets = nullptr;
  whole.world_size          = 1;
  whole.stride              = 0;
  whole.same_chunk          = true;
  return whole;
}

} /* extern "C" */
