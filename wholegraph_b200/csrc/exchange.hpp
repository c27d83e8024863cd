/* Bucket-exchange plan shared by the DISTRIBUTED gather/scatter path and the gradient path. */
#pragma once
#include "ops_internal.hpp"

namespace wm {

struct exchange_plan {
  explicit exchange_plan(wholememory_env_func_t* env);
  wholememory_dtype_t idx_dtype = WHOLEMEMORY_DT_INT64;
  int64_t n      = 0; /* indices given */
  int64_t n_send = 0; /* valid indices, grouped by owner */
  int64_t n_recv = 0; /* indices other ranks (and this one) address to me */
  std::vector<int64_t> send_counts, recv_counts; /* per rank */
  temp_buffer grouped_idx; /* [n_send] indices grouped by owner (stable) */
  temp_buffer origin;      /* [n_send] int64: position of each grouped index in the caller's array */
  temp_buffer recv_idx;    /* [n_recv] global row ids this rank must serve, grouped by requester */
  temp_buffer scratch_hist, scratch_base, scratch_totals, host_totals;
};

/* Collective.  first_row: world_size+1 partition boundaries in rows. */
void plan_exchange(exchange_plan* p,
                   wholememory_comm_t comm,
                   const void* indices,
                   wholememory_dtype_t idx_dtype,
                   int64_t n,
                   const std::vector<int64_t>& first_row,
                   cudaStream_t stream);

/* alltoallv of fixed-size rows following the plan.  to_owner: requester -> owner (send_counts out,
 * recv_counts in); otherwise the reverse direction. */
void exchange_rows(const exchange_plan& p,
                   wholememory_comm_t comm,
                   const void* send,
                   void* recv,
                   size_t row_bytes,
                   bool to_owner,
                   cudaStream_t stream);

std::vector<int64_t> handle_first_rows(wholememory_handle_t h, size_t row_stride_bytes);

}  // namespace wm
