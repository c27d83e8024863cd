/* Internal interface between the op entry points (ops.cpp, embedding.cpp, ...) and the kernels. */
#pragma once
#include "table_ref.hpp"
#include "wm_internal.hpp"

namespace wm {

/* table_ref for any tensor the ops accept: raw pointer, CONTINUOUS/CHUNKED/HOST handle, or a
 * peer-mapped DISTRIBUTED handle (internal fast path).  Throws when rows are not addressable. */
table_ref make_table_ref(wholememory_tensor_t t);
table_ref make_flat_table_ref(void* base);
/* true when every rank's shard of this handle can be dereferenced from this process */
bool handle_is_addressable(wholememory_handle_t h);

/*
 * gather : dense[i,:] = cvt(table[idx[i],:])      scatter: table[idx[i],:] = cvt(dense[i,:])
 * idx < 0 => row skipped.  indices points at element 0 (storage_offset already applied by caller? no:
 * idx_desc.storage_offset is applied here).  dense points at the allocation start; dense_desc
 * carries its storage_offset.  sms <= 0 => whole GPU.
 */
void row_move(bool gather,
              const table_ref& tref,
              const wholememory_matrix_description_t& table_desc,
              const void* indices,
              const wholememory_array_description_t& idx_desc,
              void* dense,
              const wholememory_matrix_description_t& dense_desc,
              cudaStream_t stream,
              int sms);

/* converting variants live in their own translation units (build parallelism) */
using cvt_launch_fn = void (*)(bool gather,
                               const table_ref&,
                               const row_geom&,
                               const void* indices,
                               bool idx64,
                               int64_t n,
                               char* dense,
                               int align,
                               int grid,
                               cudaStream_t stream);
cvt_launch_fn find_float_cvt(wholememory_dtype_t table_dt, wholememory_dtype_t dense_dt);
cvt_launch_fn find_int_cvt(wholememory_dtype_t table_dt, wholememory_dtype_t dense_dt);
int cvt_blocks_per_sm();

/* The two checks every 1-D operand of the graph / sampling entry points goes through in the reference, with its codes:
 * rank != 1 -> WHOLEMEMORY_INVALID_INPUT, then "cannot be viewed as an array" (last stride != 1, unknown dtype) ->
 * WHOLEMEMORY_LOGIC_ERROR (e.g. unweighted_sample_without_replacement.cpp:64-111).  The call order at each entry point
 * follows the reference's too; tests/cpp/graph_validation_diff.cpp compares it with the reference source on CPU. */
inline bool is_1d(wholememory_tensor_t t) { return wholememory_tensor_get_tensor_description(t)->dim == 1; }
inline bool views_as_array(wholememory_tensor_t t)
{
  wholememory_array_description_t a;
  wholememory_tensor_description_t d = *wholememory_tensor_get_tensor_description(t);
  return wholememory_convert_tensor_desc_to_array(&a, &d);
}

}  // namespace wm
