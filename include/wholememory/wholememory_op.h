/*
 * The hot path: row gather / scatter on a WholeMemory tensor.
 * Drop-in for reference cpp/include/wholememory/wholememory_op.h:38-78
 * (impl there: cpp/src/wholememory_ops/gather_op.cpp:23, scatter_op.cpp:23).
 * Implementation here: wholegraph_b200/csrc/ops.cpp + gather_scatter.cu (sm_100a).
 */
#pragma once
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * output[i, :] = convert(table[indices[i], :]) for every i with indices[i] >= 0; rows with a
 * negative index are left untouched.  table: 1-D/2-D WholeMemory (or raw-pointer) tensor;
 * indices: contiguous 1-D int32/int64 device array; output: device tensor of the same rank.
 * dtype pairs: float<->float family or int<->int family, converted through float for fp16/bf16.
 * Asynchronous on `stream` for CONTINUOUS/CHUNKED (and peer-mapped DISTRIBUTED) tables.
 * gather_sms: SM budget for the kernel, -1 = all.
 */
wholememory_error_code_t wholememory_gather(wholememory_tensor_t wm_tensor, wholememory_tensor_t indices_tensor,
    wholememory_tensor_t output_tensor, wholememory_env_func_t* env_fns, void* stream, int gather_sms = -1);

/* table[indices[i], :] = convert(input[i, :]); negative indices skipped; duplicate indices race
 * (last writer wins, as in the reference). */
wholememory_error_code_t wholememory_scatter(wholememory_tensor_t input_tensor, wholememory_tensor_t indices_tensor,
    wholememory_tensor_t wm_tensor, wholememory_env_func_t* env_fns, void* stream, int scatter_sms = -1);

/* allocator-plumbing self test used by the binding's unit test (reference wholememory_op.h:70-78) */
wholememory_error_code_t wholememory_env_test_op(wholememory_tensor_t input_tensor,
    wholememory_tensor_t output_fixed_tensor, void* output_variable_device_tensor_handle,
    void* output_variable_pinned_tensor_handle, void* output_variable_host_tensor_handle,
    int64_t output_variable_entry_count, wholememory_env_func_t* env_fns, void* stream);

#ifdef __cplusplus
}
#endif
