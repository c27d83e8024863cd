/*
 * WholeMemory embedding objects: padded-row embedding table + sparse optimizer.
 * Drop-in for reference cpp/include/wholememory/embedding.h:44-244 (C API impl there:
 * cpp/src/wholememory/embedding.cpp:900-1152, optimizers cpp/src/wholememory/embedding_optimizer.cpp).
 * Implementation here: wholegraph_b200/csrc/embedding.cpp + sparse_optimizer.cu.
 *
 * Scope: the non-cached embedding (gather == wholememory_gather on the padded table; gradient
 * apply == owner exchange + duplicate merge + ONE fused optimizer kernel).  Cache policies are
 * accepted as opaque objects, but creating an embedding WITH a cache policy returns
 * WHOLEMEMORY_NOT_IMPLEMENTED: the device-cache-for-host design exists to hide PCIe latency
 * and a B200 box keeps tables in HBM (DESIGN.md "out of scope").
 */
#pragma once
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wholememory_embedding_cache_policy_* wholememory_embedding_cache_policy_t;
typedef struct wholememory_embedding_optimizer_* wholememory_embedding_optimizer_t;
typedef struct wholememory_embedding_* wholememory_embedding_t;

/* reference embedding.h:44-48 */
enum wholememory_access_type_t {
  WHOLEMEMORY_AT_NONE = 0,
  WHOLEMEMORY_AT_READONLY,
  WHOLEMEMORY_AT_READWRITE,
};

/* reference embedding.h:53-59 */
enum wholememory_optimizer_type_t {
  WHOLEMEMORY_OPT_NONE = 0,
  WHOLEMEMORY_OPT_SGD,
  WHOLEMEMORY_OPT_LAZY_ADAM,
  WHOLEMEMORY_OPT_RMSPROP,
  WHOLEMEMORY_OPT_ADAGRAD,
};

wholememory_error_code_t wholememory_create_embedding_optimizer(wholememory_embedding_optimizer_t* optimizer,
    wholememory_optimizer_type_t optimizer_type);
/* parameter_name in {"weight_decay","epsilon","beta1","beta2","adam_w","alpha"}; value -> float */
wholememory_error_code_t wholememory_optimizer_set_parameter(wholememory_embedding_optimizer_t optimizer,
    const char* parameter_name, void* value);
void wholememory_destroy_embedding_optimizer(wholememory_embedding_optimizer_t optimizer);

/* cache_ratio must lie in [1/512, 1] */
wholememory_error_code_t wholememory_create_embedding_cache_policy(
    wholememory_embedding_cache_policy_t* cache_policy, wholememory_comm_t cache_level_comm,
    wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location,
    wholememory_access_type_t access_type, float cache_ratio);
wholememory_error_code_t wholememory_destroy_embedding_cache_policy(
    wholememory_embedding_cache_policy_t cache_policy);

/* Collective.  The stored row stride is padded to a multiple of 16 bytes; the tensor returned by
 * wholememory_embedding_get_embedding_tensor is the [N, D] view of it. */
wholememory_error_code_t wholememory_create_embedding(wholememory_embedding_t* embedding,
    wholememory_tensor_description_t* embedding_tensor_description, wholememory_comm_t comm,
    wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location,
    wholememory_embedding_cache_policy_t cache_policy, size_t* embedding_entry_partition = nullptr,
    int user_defined_sms = -1, int round_robin_size = 0);
wholememory_error_code_t wholememory_destroy_embedding(wholememory_embedding_t embedding);
wholememory_tensor_t wholememory_embedding_get_embedding_tensor(wholememory_embedding_t embedding);
/* at most once, before training; fp32 embeddings only; allocates the optimizer state tables */
wholememory_error_code_t wholememory_embedding_set_optimizer(wholememory_embedding_t embedding,
    wholememory_embedding_optimizer_t optimizer);

wholememory_error_code_t wholememory_embedding_gather(wholememory_embedding_t embedding,
    wholememory_tensor_t indices, wholememory_tensor_t output, bool adjust_cache, wholememory_env_func_t* env_fns,
    int64_t stream_int);
/* Collective.  indices/grads are this rank's (row id, fp32 gradient row) pairs; gradients of
 * duplicate ids (from any rank) are summed, then the owner applies ONE optimizer step per row. */
wholememory_error_code_t wholememory_embedding_gather_gradient_apply(wholememory_embedding_t embedding,
    wholememory_tensor_t indices, wholememory_tensor_t grads, bool adjust_cache, float lr,
    wholememory_env_func_t* env_fns, int64_t stream_int);

/* nullptr-terminated list, e.g. {"m","v","beta12t",nullptr} for LazyAdam */
const char* const* wholememory_embedding_get_optimizer_state_names(
  wholememory_embedding_t embedding);
wholememory_tensor_t wholememory_embedding_get_optimizer_state(wholememory_embedding_t embedding, const char* name);

/* no cache in this build: both succeed as no-ops */
wholememory_error_code_t wholememory_embedding_writeback_cache(wholememory_embedding_t embedding, int64_t stream_int);
wholememory_error_code_t wholememory_embedding_drop_all_cache(wholememory_embedding_t embedding, int64_t stream_int);

#ifdef __cplusplus
}
#endif
