/*
 * C-ABI entry points of the hot path: wholememory_gather / wholememory_scatter.
 * Validation order and error codes follow reference cpp/src/wholememory_ops/gather_op.cpp:23-131
 * and scatter_op.cpp:23-110; dispatch differs:
 *   - CONTINUOUS / CHUNKED / HOST / raw pointer  -> one peer-load (peer-store) kernel
 *   - DISTRIBUTED, peer-mapped (every HGX box)    -> the same kernel (no collective)
 *   - DISTRIBUTED, not addressable                -> bucket exchange over NCCL (exchange.cu)
 * The reference's "sort indices first for HOST tables with rows <= 512 B" branch
 * (gather_op.cpp:116-120) only reorders reads; results are identical and it is not rebuilt.
 */
#include "ops_internal.hpp"

namespace wm {

bool handle_is_addressable(wholememory_handle_t h)
{
  /* WG_FORCE_EXCHANGE=1: route DISTRIBUTED ops through the NCCL bucket exchange even when the shards
   * are peer-mapped (exercises the no-NVLink path on an NVSwitch box and on 1 rank) */
  static const bool force_exchange = [] {
    const char* v = getenv("WG_FORCE_EXCHANGE");
    return v != nullptr && v[0] != '\0' && v[0] != '0';
  }();
  if (force_exchange && h->type == WHOLEMEMORY_MT_DISTRIBUTED) return false;
  return h->peer_mapped;
}

table_ref make_table_ref(wholememory_tensor_t t)
{
  if (!t->is_wm) {
    WM_EXPECT(t->storage != nullptr, WHOLEMEMORY_INVALID_INPUT, "tensor has no storage");
    return make_flat_table_ref(t->storage);
  }
  wholememory_handle_t h = t->handle;
  WM_EXPECT(h->peer_mapped, WHOLEMEMORY_LOGIC_ERROR, "WholeMemory handle is not addressable from this rank");
  const int ws         = h->comm->world_size;
  const int has_remote = (ws > 1 || h->location == WHOLEMEMORY_ML_HOST) ? 1 : 0;
  if (h->flat_base != nullptr) {
    table_ref f  = make_flat_table_ref(h->flat_base);
    f.has_remote = has_remote;
    return f;
  }
  table_ref r{};
  r.nranks      = ws;
  r.has_remote  = has_remote;
  r.chunk_bytes = h->chunk_stride;
  if (ws <= kMaxInlineRanks) {
    r.mode = h->regular ? table_ref::CHUNK_REGULAR : table_ref::CHUNK_IRREGULAR;
    for (int i = 0; i < ws; ++i) {
      r.base[i]       = static_cast<char*>(h->rank_base[i]);
      r.first_byte[i] = h->part_offsets[i];
    }
    r.first_byte[ws] = h->part_offsets[ws];
  } else {
    WM_EXPECT(h->d_chunk_table != nullptr, WHOLEMEMORY_NOT_SUPPORTED, "more than %d ranks need a CHUNKED table", kMaxInlineRanks);
    r.mode           = h->regular ? table_ref::DEVTAB_REGULAR : table_ref::DEVTAB_IRREGULAR;
    r.dev_bases      = reinterpret_cast<char* const*>(h->d_chunk_table);
    r.dev_first_byte = reinterpret_cast<const uint64_t*>(h->d_offsets);
  }
  return r;
}

/* exchange.cu */
void gather_by_exchange(wholememory_handle_t h,
                        const wholememory_matrix_description_t& table_desc,
                        const void* indices,
                        const wholememory_array_description_t& idx_desc,
                        void* output,
                        const wholememory_matrix_description_t& out_desc,
                        wholememory_env_func_t* env,
                        cudaStream_t stream,
                        int sms);
void scatter_by_exchange(const void* input,
                         const wholememory_matrix_description_t& in_desc,
                         const void* indices,
                         const wholememory_array_description_t& idx_desc,
                         wholememory_handle_t h,
                         const wholememory_matrix_description_t& table_desc,
                         wholememory_env_func_t* env,
                         cudaStream_t stream,
                         int sms);

namespace {

struct op_args {
  wholememory_matrix_description_t table, dense;
  wholememory_array_description_t idx;
  void* idx_ptr;
  void* dense_ptr;
};

/* Shared descriptor checks, in the reference's order and with its error codes (gather_op.cpp:38-82, scatter_op.cpp:38-79;
 * pinned by tests/cpp/ops_validation_diff.cpp against those files compiled for the CPU).  `what` names the dense operand.
 * Two quirks of the reference are kept on purpose:
 *  - a 1-D table is unsqueezed to [N, 1] BEFORE the rank comparison, so the dense operand it expects for a 1-D table is
 *    the 2-D [n, 1] one.  (A 1-D dense operand for a 1-D table, which the reference refuses, is accepted here as well.)
 *  - a table that cannot be viewed as a matrix is WHOLEMEMORY_LOGIC_ERROR for gather but WHOLEMEMORY_INVALID_INPUT for scatter. */
wholememory_error_code_t prepare(wholememory_tensor_t table,
                                 wholememory_tensor_t indices,
                                 wholememory_tensor_t dense,
                                 const char* what,
                                 bool is_gather,
                                 op_args* a)
{
  WM_REQUIRE_LIVE(table);
  WM_REQUIRE_LIVE(indices);
  WM_REQUIRE_LIVE(dense);
  const wholememory_error_code_t bad_table = is_gather ? WHOLEMEMORY_LOGIC_ERROR : WHOLEMEMORY_INVALID_INPUT;
  wholememory_tensor_description_t td = *wholememory_tensor_get_tensor_description(table);
  if (td.dim != 1 && td.dim != 2) {
    WM_ERROR("wholememory_tensor should be 1D or 2D tensor.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  const int table_dim = td.dim;
  if (td.dim == 1 && !wholememory_unsqueeze_tensor(&td, 1)) return bad_table;
  if (!wholememory_convert_tensor_desc_to_matrix(&a->table, &td)) {
    WM_ERROR("wholememory_tensor cannot be viewed as a matrix.");
    return bad_table;
  }
  if (wholememory_tensor_get_tensor_description(indices)->dim != 1) {
    WM_ERROR("indices tensor should be 1D tensor");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  wholememory_tensor_description_t dd = *wholememory_tensor_get_tensor_description(dense);
  if (dd.dim != 2 && !(dd.dim == 1 && table_dim == 1)) {
    WM_ERROR("%s tensor should be same dim as wholememory_tensor.", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (dd.dim == 1 && !wholememory_unsqueeze_tensor(&dd, 1)) return WHOLEMEMORY_LOGIC_ERROR;
  if (!wholememory_convert_tensor_desc_to_array(&a->idx, wholememory_tensor_get_tensor_description(indices))) {
    WM_ERROR("Convert indices tensor to array failed.");
    return WHOLEMEMORY_INVALID_INPUT;
  }
  if (!wholememory_convert_tensor_desc_to_matrix(&a->dense, &dd)) {
    WM_ERROR("Convert %s tensor to matrix failed.", what);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  /* indices / dense operands are caller memory: raw pointer tensors (or CONTINUOUS handles).
   * get_data_pointer applies storage_offset, so zero the offsets we pass on. */
  a->idx_ptr   = wholememory_tensor_get_data_pointer(indices);
  a->dense_ptr = wholememory_tensor_get_data_pointer(dense);
  a->idx.storage_offset   = 0;
  a->dense.storage_offset = 0;
  return WHOLEMEMORY_SUCCESS;
}

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t wholememory_gather(wholememory_tensor_t wholememory_tensor,
                                            wholememory_tensor_t indices_tensor,
                                            wholememory_tensor_t output_tensor,
                                            wholememory_env_func_t* p_env_fns,
                                            void* stream,
                                            int gather_sms)
{
  return wm::guarded("wholememory_gather", [&]() -> wholememory_error_code_t {
    wm::op_args a;
    auto rc = wm::prepare(wholememory_tensor, indices_tensor, output_tensor, "output", true, &a);
    if (rc != WHOLEMEMORY_SUCCESS) return rc;
    auto s = static_cast<cudaStream_t>(stream);
    if (wholememory_tensor->is_wm && !wm::handle_is_addressable(wholememory_tensor->handle)) {
      wm::gather_by_exchange(wholememory_tensor->handle, a.table, a.idx_ptr, a.idx, a.dense_ptr, a.dense, p_env_fns, s, gather_sms);
      return WHOLEMEMORY_SUCCESS;
    }
    wm::row_move(true, wm::make_table_ref(wholememory_tensor), a.table, a.idx_ptr, a.idx, a.dense_ptr, a.dense, s, gather_sms);
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_scatter(wholememory_tensor_t input_tensor,
                                             wholememory_tensor_t indices_tensor,
                                             wholememory_tensor_t wholememory_tensor,
                                             wholememory_env_func_t* p_env_fns,
                                             void* stream,
                                             int scatter_sms)
{
  return wm::guarded("wholememory_scatter", [&]() -> wholememory_error_code_t {
    wm::op_args a;
    auto rc = wm::prepare(wholememory_tensor, indices_tensor, input_tensor, "input", false, &a);
    if (rc != WHOLEMEMORY_SUCCESS) return rc;
    auto s = static_cast<cudaStream_t>(stream);
    if (wholememory_tensor->is_wm && !wm::handle_is_addressable(wholememory_tensor->handle)) {
      wm::scatter_by_exchange(a.dense_ptr, a.dense, a.idx_ptr, a.idx, wholememory_tensor->handle, a.table, p_env_fns, s, scatter_sms);
      return WHOLEMEMORY_SUCCESS;
    }
    wm::row_move(false, wm::make_table_ref(wholememory_tensor), a.table, a.idx_ptr, a.idx, a.dense_ptr, a.dense, s, scatter_sms);
    return WHOLEMEMORY_SUCCESS;
  });
}

} /* extern "C" */
