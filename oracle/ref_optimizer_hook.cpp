/*
 * TEST INFRASTRUCTURE (oracle/_ref build only): one extern "C" entry that drives the REFERENCE's own
 * duplicate-gradient merge and sparse-optimizer kernels, so the GPU tests can compare this repo's fused
 * merge+update kernel with the reference binary, not only with a CPU model.
 *
 * It performs the owner-local tail of the reference's embedding_base::gather_gradient_apply
 * (cpp/src/wholememory/embedding.cpp:257-316) at world_size 1, where the preceding bucket/exchange is the identity
 * up to a stable sort:  dedup_indice_and_gradients (functions/exchange_embeddings_nccl_func.cu:180-206)  ->
 * make_tensor_from_pointer x2  ->  {sgd,lazy_adam,ada_grad,rms_prop}_optimizer_step
 * (functions/embedding_optimizer_func.cu, built from the reference tree with the RAFT stand-in
 * oracle/ref_shim/raft/matrix/detail/select_k-inl.cuh; argument order as embedding_optimizer.cpp:129-140,
 * :259-277, :341-359, :450-468).  The cache arguments are null / coverage 0 (non-cached embedding).
 * Nothing here is linked into libwholegraph.so.
 */
#include <cuda_runtime_api.h>

#include <wholememory/embedding.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory_tensor.h>

#include "wholememory/env_func_ptrs.hpp"
#include "wholememory_ops/functions/embedding_optimizer_func.h"
#include "wholememory_ops/functions/exchange_embeddings_nccl_func.h"

extern "C" wholememory_error_code_t wgref_dedup_and_optimizer_step(int optimizer_type,
                                                                   wholememory_tensor_t indices,
                                                                   wholememory_tensor_t grads,
                                                                   wholememory_tensor_t local_embedding,
                                                                   wholememory_tensor_t per_element_local_state,
                                                                   wholememory_tensor_t per_embedding_local_state,
                                                                   int64_t local_entry_offset,
                                                                   float weight_decay,
                                                                   float epsilon,
                                                                   float beta1,
                                                                   float beta2,
                                                                   int adam_w,
                                                                   float alpha,
                                                                   float lr,
                                                                   wholememory_env_func_t* env,
                                                                   void* stream_ptr,
                                                                   int64_t* deduped_count_out)
{
  auto stream = static_cast<cudaStream_t>(stream_ptr);
  if (env == nullptr) env = wholememory::get_default_env_func();
  auto* idx_desc  = wholememory_tensor_get_tensor_description(indices);
  auto* grad_desc = wholememory_tensor_get_tensor_description(grads);
  wholememory_array_description_t idx_arr;
  wholememory_matrix_description_t grad_mat;
  if (!wholememory_convert_tensor_desc_to_array(&idx_arr, idx_desc)) return WHOLEMEMORY_INVALID_INPUT;
  if (!wholememory_convert_tensor_desc_to_matrix(&grad_mat, grad_desc)) return WHOLEMEMORY_INVALID_INPUT;
  const int64_t n   = idx_arr.size;
  const int64_t dim = grad_mat.sizes[1];
  if (grad_mat.stride != dim) return WHOLEMEMORY_INVALID_INPUT; /* the reference passes a packed receive buffer */

  void* dedup_indice = nullptr;
  float* dedup_grads = nullptr;
  size_t idx_bytes   = (size_t)(n > 0 ? n : 1) * wholememory_dtype_get_element_size(idx_arr.dtype);
  if (cudaMalloc(&dedup_indice, idx_bytes) != cudaSuccess) return WHOLEMEMORY_OUT_OF_MEMORY;
  if (cudaMalloc(reinterpret_cast<void**>(&dedup_grads), (size_t)(n > 0 ? n : 1) * dim * sizeof(float)) != cudaSuccess) {
    cudaFree(dedup_indice);
    return WHOLEMEMORY_OUT_OF_MEMORY;
  }

  int64_t deduped = wholememory_ops::dedup_indice_and_gradients(wholememory_tensor_get_data_pointer(indices),
                                                                idx_arr,
                                                                static_cast<const float*>(wholememory_tensor_get_data_pointer(grads)),
                                                                grad_mat,
                                                                dedup_indice,
                                                                dedup_grads,
                                                                env,
                                                                stream);
  if (deduped_count_out != nullptr) *deduped_count_out = deduped;

  wholememory_tensor_t dedup_idx_tensor = nullptr, dedup_grad_tensor = nullptr;
  wholememory_tensor_description_t di = *idx_desc;
  di.sizes[0]                         = deduped;
  wholememory_tensor_description_t dg = *grad_desc;
  dg.sizes[0]                         = deduped;
  dg.strides[0]                       = dg.sizes[1];
  wholememory_error_code_t rc         = wholememory_make_tensor_from_pointer(&dedup_idx_tensor, dedup_indice, &di);
  if (rc == WHOLEMEMORY_SUCCESS) rc = wholememory_make_tensor_from_pointer(&dedup_grad_tensor, dedup_grads, &dg);
  if (rc == WHOLEMEMORY_SUCCESS) {
    switch (optimizer_type) {
      case WHOLEMEMORY_OPT_SGD:
        rc = wholememory_ops::sgd_optimizer_step(dedup_idx_tensor, dedup_grad_tensor, local_embedding, nullptr, nullptr,
                                                 local_entry_offset, 0, weight_decay, lr, stream);
        break;
      case WHOLEMEMORY_OPT_LAZY_ADAM:
        rc = wholememory_ops::lazy_adam_optimizer_step(dedup_idx_tensor, dedup_grad_tensor, local_embedding, nullptr, nullptr,
                                                       per_element_local_state, nullptr, nullptr, per_embedding_local_state,
                                                       local_entry_offset, 0, weight_decay, epsilon, beta1, beta2, adam_w != 0,
                                                       lr, stream);
        break;
      case WHOLEMEMORY_OPT_ADAGRAD:
        rc = wholememory_ops::ada_grad_optimizer_step(dedup_idx_tensor, dedup_grad_tensor, local_embedding, nullptr, nullptr,
                                                      per_element_local_state, nullptr, nullptr, local_entry_offset, 0,
                                                      weight_decay, epsilon, lr, stream);
        break;
      case WHOLEMEMORY_OPT_RMSPROP:
        rc = wholememory_ops::rms_prop_optimizer_step(dedup_idx_tensor, dedup_grad_tensor, local_embedding, nullptr, nullptr,
                                                      per_element_local_state, nullptr, nullptr, local_entry_offset, 0,
                                                      weight_decay, epsilon, alpha, lr, stream);
        break;
      default: rc = WHOLEMEMORY_INVALID_INPUT; break;
    }
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess && rc == WHOLEMEMORY_SUCCESS) rc = WHOLEMEMORY_CUDA_ERROR;
  if (dedup_idx_tensor) wholememory_destroy_tensor(dedup_idx_tensor);
  if (dedup_grad_tensor) wholememory_destroy_tensor(dedup_grad_tensor);
  cudaFree(dedup_indice);
  cudaFree(dedup_grads);
  return rc;
}
