/* Interface of the fused duplicate-merge + sparse optimizer kernel (sparse_optimizer.cu). */
#pragma once
#include "wm_internal.hpp"

namespace wm {

/* hyper-parameters, defaults as reference cpp/src/wholememory/embedding_optimizer.cpp:170-174, :404 */
struct optimizer_params {
  float weight_decay = 0.0f;
  float epsilon      = 1e-8f;
  float beta1        = 0.9f;
  float beta2        = 0.999f;
  int adam_w         = 0;
  float alpha        = 0.99f;
};

/* this rank's rows of the embedding and of its optimizer state */
struct optimizer_rows {
  float* w;                /* local embedding rows */
  int64_t w_stride;        /* padded row stride in floats (multiple of 4) */
  float* state;            /* per-element state rows: [m | v] (LazyAdam), [state_sum] (AdaGrad), [v] (RMSProp); null for SGD */
  int64_t state_stride;    /* floats per state row */
  float* b12;              /* LazyAdam: [local_rows, 2] beta1^t, beta2^t */
  int64_t local_row_start; /* global id of local row 0 */
  int64_t local_rows;
  int dim;                 /* embedding width */
};

/*
 * For every distinct id in indices (global row ids, duplicates allowed, arrival order significant
 * only for the fp32 summation order): g = sum of its gradient rows; apply one optimizer step to the
 * local row.  ids outside [local_row_start, local_row_start+local_rows) are ignored.
 */
void merge_and_update_rows(int optimizer_type,
                           const void* indices,
                           wholememory_dtype_t idx_dtype,
                           int64_t n,
                           const float* grads,
                           int64_t grad_stride,
                           const optimizer_rows& rows,
                           const optimizer_params& params,
                           float lr,
                           int64_t total_rows,
                           bool may_have_negative, /* ids come straight from the caller (not filtered by the exchange) */
                           wholememory_env_func_t* env,
                           cudaStream_t stream);

void fill_float(float* p, float value, int64_t n, cudaStream_t stream);

}  // namespace wm
