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

}  // namespace wm
