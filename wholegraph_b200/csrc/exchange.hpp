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

/* Local part of the plan: grouped_idx / origin / send_counts / n_send.  No communication.
 * first_row: world_size+1 partition boundaries in rows. */
void partition_by_owner(exchange_plan* p,
                        wholememory_comm_t comm,
                        const void* indices,
                        wholememory_dtype_t idx_dtype,
                        int64_t n,
                        const std::vector<int64_t>& first_row,
                        cudaStream_t stream);

/* Collective: partition_by_owner + count all-to-all + shipping the grouped indices to their owners (NCCL). */
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

/*
 * Peer-store push of (row id, fp32 row) pairs to the owners' staging areas (peer_push.cu): the gradient exchange on an
 * NVSwitch box.  The stage is a CHUNKED/DEVICE WholeMemory allocation every rank maps; it is double-buffered so that the
 * owner's update kernel of step k may still run while step k+1's rows arrive.
 */
struct push_stage {
  wholememory_handle_t h = nullptr;
  int64_t cap_rows       = 0; /* rows per rank and buffer */
  int64_t dim            = 0;
  int flip               = 0;
};

/* Collective.  `p` must come from partition_by_owner.  On return every row addressed to this rank sits in its stage, in
 * (sender rank, sender order) order - the order the NCCL exchange delivers - and is visible to `stream`:
 * *ids = [n_recv] int64 global row ids, *rows = [n_recv, dim] fp32.  Returns n_recv. */
int64_t push_rows_to_owners(push_stage* st,
                            wholememory_comm_t comm,
                            const exchange_plan& p,
                            const float* rows_in,
                            int64_t row_stride,
                            int64_t dim,
                            cudaStream_t stream,
                            const int64_t** ids,
                            const float** rows);
void destroy_push_stage(push_stage* st);

}  // namespace wm
