/*
 * WholeMemory tensor objects: strided 1-D/2-D views over a handle or a raw pointer.
 * Behaviour follows reference cpp/src/wholememory/wholememory_tensor.cpp:51-469.
 */
#include "wm_internal.hpp"

#include <algorithm>
#include <atomic>

namespace {

std::atomic<int64_t> g_live_tensors{0};

bool known_dtype(wholememory_dtype_t d) { return d > WHOLEMEMORY_DT_UNKNOWN && d < WHOLEMEMORY_DT_COUNT; }

wholememory_tensor_t new_tensor()
{
  auto* t = new wholememory_tensor_();
  t->root = t;
  ++g_live_tensors;
  wm::obj_register(wm::OBJ_TENSOR, t);
  return t;
}

/* bytes of one partition entry (row) of the ROOT allocation a tensor lives in */
size_t entry_bytes(wholememory_tensor_t t)
{
  const auto& rd = t->root->desc;
  size_t elems   = rd.dim == 2 ? (size_t)rd.strides[0] : 1;
  return elems * wholememory_dtype_get_element_size(t->desc.dtype);
}

}  // namespace

extern "C" {

int64_t get_wholememory_tensor_count() { return g_live_tensors.load(); }

wholememory_error_code_t wholememory_create_tensor(wholememory_tensor_t* out,
                                                   wholememory_tensor_description_t* desc,
                                                   wholememory_comm_t comm,
                                                   wholememory_memory_type_t memory_type,
                                                   wholememory_memory_location_t memory_location,
                                                   size_t* tensor_entry_partition)
{
  if (out == nullptr || desc == nullptr) {
    WM_ERROR("wholememory_create_tensor: null argument");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (desc->dim < 1 || desc->dim > 2 || desc->storage_offset != 0 || desc->strides[desc->dim - 1] != 1 ||
      !known_dtype(desc->dtype)) {
    WM_ERROR("wholememory_create_tensor: need a 1-D/2-D, offset-0, unit-inner-stride, typed description "
             "(dim=%d offset=%ld)", desc->dim, (long)desc->storage_offset);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  size_t esize       = wholememory_dtype_get_element_size(desc->dtype);
  size_t bytes       = (size_t)wholememory_get_memory_element_count_from_tensor(desc) * esize;
  size_t granularity = esize * (size_t)desc->strides[0]; /* one row (reference :85-86) */
  wholememory_handle_t h = nullptr;
  auto rc = wholememory_malloc(&h, bytes, comm, memory_type, memory_location, granularity, tensor_entry_partition);
  if (rc != WHOLEMEMORY_SUCCESS) return rc;
  auto* t       = new_tensor();
  t->handle     = h;
  t->desc       = *desc;
  t->is_wm      = true;
  t->own_handle = true;
  *out          = t;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_destroy_tensor(wholememory_tensor_t t)
{
  WM_REQUIRE_KNOWN(wm::OBJ_TENSOR, t);
  /* a handle that died with its communicator (wholememory_finalize) is already gone: only the tensor object is left */
  if (t->own_handle && t->is_wm && wm::live(t->handle)) WHOLEMEMORY_RETURN_ON_FAIL(wholememory_free(t->handle));
  wm::obj_unregister(wm::OBJ_TENSOR, t);
  --g_live_tensors;
  delete t;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_make_tensor_from_pointer(wholememory_tensor_t* out,
                                                              void* storage_ptr,
                                                              wholememory_tensor_description_t* desc)
{
  if (out == nullptr || desc == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  /* empty tensors (null storage or dim 0) are accepted unchecked, like the reference (:129-141) */
  bool empty = storage_ptr == nullptr || desc->dim == 0;
  if (!empty) {
    if (desc->dim < 0 || desc->dim > WHOLEMEMORY_MAX_TENSOR_DIM || desc->strides[desc->dim - 1] != 1 ||
        !known_dtype(desc->dtype)) {
      WM_ERROR("wholememory_make_tensor_from_pointer: bad description (dim=%d dtype=%d)", desc->dim, (int)desc->dtype);
      return WHOLEMEMORY_INVALID_INPUT;
    }
  }
  auto* t    = new_tensor();
  t->storage = storage_ptr;
  t->desc    = *desc;
  *out       = t;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_make_tensor_from_handle(wholememory_tensor_t* out,
                                                             wholememory_handle_t handle,
                                                             wholememory_tensor_description_t* desc)
{
  if (out == nullptr || desc == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(handle);
  if (desc->dim < 1 || desc->dim > 2 || desc->strides[desc->dim - 1] != 1 || !known_dtype(desc->dtype)) {
    WM_ERROR("wholememory_make_tensor_from_handle: bad description (dim=%d dtype=%d)", desc->dim, (int)desc->dtype);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  auto* t   = new_tensor();
  t->handle = handle;
  t->desc   = *desc;
  t->is_wm  = true;
  *out      = t;
  return WHOLEMEMORY_SUCCESS;
}

/* pointer / bool getters have no error channel: an unknown tensor reads as "no handle" / nullptr */
bool wholememory_tensor_has_handle(wholememory_tensor_t t) { return wm::obj_known(wm::OBJ_TENSOR, t) && t->is_wm; }

wholememory_handle_t wholememory_tensor_get_memory_handle(wholememory_tensor_t t)
{
  return wm::live(t) && t->is_wm ? t->handle : nullptr;
}

wholememory_tensor_description_t* wholememory_tensor_get_tensor_description(wholememory_tensor_t t)
{
  return wm::obj_known(wm::OBJ_TENSOR, t) ? &t->desc : nullptr;
}

wholememory_error_code_t wholememory_tensor_get_global_reference(wholememory_tensor_t t, wholememory_gref_t* gref)
{
  if (gref == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (t->is_wm) return wholememory_get_global_reference(gref, t->handle);
  *gref = wholememory_create_continuous_global_reference(t->storage);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_map_local_tensor(wholememory_tensor_t t, wholememory_tensor_t* local)
{
  if (local == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (!t->is_wm) return WHOLEMEMORY_INVALID_VALUE;
  const auto& d = t->desc;
  /* the view may drop trailing rows/columns but must start at row 0 (reference :238-245) */
  if (d.dim != 1 && d.dim != 2) return WHOLEMEMORY_INVALID_VALUE;
  if (d.dim == 1 && d.storage_offset != 0) return WHOLEMEMORY_INVALID_VALUE;
  if (d.dim == 2 && d.storage_offset + d.sizes[1] > d.strides[0]) return WHOLEMEMORY_INVALID_VALUE;
  void* ptr     = nullptr;
  size_t bytes  = 0, offset = 0;
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_get_local_memory(&ptr, &bytes, &offset, t->handle));
  size_t esize = wholememory_dtype_get_element_size(d.dtype);
  size_t gran  = d.dim == 1 ? esize : esize * (size_t)d.strides[0];
  size_t limit = (size_t)d.sizes[0] * gran; /* bytes covered by the (possibly row-truncated) view */
  bytes        = offset >= limit ? 0 : std::min(bytes, limit - offset);
  if (bytes % gran != 0) return WHOLEMEMORY_LOGIC_ERROR;
  wholememory_tensor_description_t ld = d;
  ld.sizes[0]                         = (int64_t)(bytes / gran);
  return wholememory_make_tensor_from_pointer(local, ptr, &ld);
}

void* wholememory_tensor_get_data_pointer(wholememory_tensor_t t)
{
  char* base = nullptr;
  if (!wm::live(t)) return nullptr;
  if (t->is_wm) {
    if (wholememory_get_memory_type(t->handle) != WHOLEMEMORY_MT_CONTINUOUS) return nullptr;
    if (wholememory_get_global_pointer(reinterpret_cast<void**>(&base), t->handle) != WHOLEMEMORY_SUCCESS) return nullptr;
  } else {
    base = static_cast<char*>(t->storage);
  }
  return base + wholememory_dtype_get_element_size(t->desc.dtype) * t->desc.storage_offset;
}

wholememory_error_code_t wholememory_tensor_get_entry_offsets(size_t* entry_offsets, wholememory_tensor_t t)
{
  if (entry_offsets == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (!t->is_wm) {
    entry_offsets[0] = 0;
    entry_offsets[1] = (size_t)t->root->desc.sizes[0];
    return WHOLEMEMORY_SUCCESS;
  }
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_get_rank_partition_offsets(entry_offsets, t->handle));
  size_t eb = entry_bytes(t);
  for (int r = 0; r <= t->handle->comm->world_size; ++r) {
    if (entry_offsets[r] % eb != 0) return WHOLEMEMORY_LOGIC_ERROR;
    entry_offsets[r] /= eb;
  }
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_get_entry_partition_sizes(size_t* entry_partition, wholememory_tensor_t t)
{
  if (entry_partition == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (!t->is_wm) {
    entry_partition[0] = (size_t)t->root->desc.sizes[0];
    return WHOLEMEMORY_SUCCESS;
  }
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_get_rank_partition_sizes(entry_partition, t->handle));
  size_t eb = entry_bytes(t);
  for (int r = 0; r < t->handle->comm->world_size; ++r) {
    if (entry_partition[r] % eb != 0) return WHOLEMEMORY_LOGIC_ERROR;
    entry_partition[r] /= eb;
  }
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_get_local_entry_count(size_t* local_entry_count, wholememory_tensor_t t)
{
  if (local_entry_count == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (!t->is_wm) {
    *local_entry_count = (size_t)t->root->desc.sizes[0];
    return WHOLEMEMORY_SUCCESS;
  }
  size_t bytes = 0;
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_get_local_size(&bytes, t->handle));
  size_t eb = entry_bytes(t);
  if (bytes % eb != 0) return WHOLEMEMORY_LOGIC_ERROR;
  *local_entry_count = bytes / eb;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_get_local_entry_start(size_t* local_entry_start, wholememory_tensor_t t)
{
  if (local_entry_start == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  if (!t->is_wm) {
    *local_entry_start = 0;
    return WHOLEMEMORY_SUCCESS;
  }
  size_t off = 0;
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_get_local_offset(&off, t->handle));
  size_t eb = entry_bytes(t);
  if (off % eb != 0) return WHOLEMEMORY_LOGIC_ERROR;
  *local_entry_start = off / eb;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_tensor_get_subtensor(wholememory_tensor_t t,
                                                          int64_t* starts,
                                                          int64_t* ends,
                                                          wholememory_tensor_t* sub)
{
  if (starts == nullptr || ends == nullptr || sub == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_LIVE(t);
  const auto& d = t->desc;
  if (d.dim > 2) return WHOLEMEMORY_NOT_IMPLEMENTED;
  wholememory_tensor_description_t nd = d;
  for (int i = 0; i < d.dim; ++i) {
    int64_t b = starts[i] == -1 ? 0 : starts[i];
    int64_t e = ends[i] == -1 ? d.sizes[i] : ends[i];
    if (e <= b || b >= d.sizes[i] || e <= 0) return WHOLEMEMORY_INVALID_INPUT;
    nd.storage_offset += d.strides[i] * b;
    nd.sizes[i] = e - b;
  }
  auto* s       = new_tensor();
  s->handle     = t->handle;
  s->storage    = t->storage;
  s->is_wm      = t->is_wm;
  s->own_handle = false;
  s->root       = t->root;
  s->desc       = nd;
  *sub          = s;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_tensor_t wholememory_tensor_get_root(wholememory_tensor_t t) { return wm::live(t) ? t->root : nullptr; }

} /* extern "C" */
