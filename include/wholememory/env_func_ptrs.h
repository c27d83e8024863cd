/*
 * Allocation callbacks ("env functions") through which ops obtain temporary and output
 * buffers from the embedding framework (torch).  Struct layouts and the call protocol
 * create_ctx -> malloc(desc, kind, ctx) -> free(ctx) -> destroy_ctx are ABI and match
 * reference cpp/include/wholememory/env_func_ptrs.h:34-75.
 */
#pragma once
#include <cuda_runtime_api.h>
#include <wholememory/tensor_description.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wholememory_memory_allocation_type_t {
  WHOLEMEMORY_MA_NONE = 0,
  WHOLEMEMORY_MA_DEVICE,
  WHOLEMEMORY_MA_HOST,
  WHOLEMEMORY_MA_PINNED,
};

typedef void (*wholememory_create_memory_context_func_t)(void** memory_context, void* global_context);
typedef void (*wholememory_destroy_memory_context_func_t)(void* memory_context, void* global_context);
typedef void* (*wholememory_malloc_func_t)(wholememory_tensor_description_t* desc,
                                           wholememory_memory_allocation_type_t memory_allocation_type,
                                           void* memory_context,
                                           void* global_context);
typedef void (*wholememory_free_func_t)(void* memory_context, void* global_context);

struct wholememory_temp_memory_func_t {
  wholememory_create_memory_context_func_t create_memory_context_fn;
  wholememory_destroy_memory_context_func_t destroy_memory_context_fn;
  wholememory_malloc_func_t malloc_fn;
  wholememory_free_func_t free_fn;
  void* global_context;
};

struct wholememory_output_memory_func_t {
  wholememory_malloc_func_t malloc_fn;
  wholememory_free_func_t free_fn;
  void* global_context;
};

struct wholememory_env_func_t {
  wholememory_temp_memory_func_t temporary_fns; /* scratch, freed before the op returns */
  wholememory_output_memory_func_t output_fns;  /* variable-size op outputs, owned by the caller's context */
};

/* cached cudaDeviceProp of device dev_id (-1 = current device) */
cudaDeviceProp* get_device_prop(int dev_id);

#ifdef __cplusplus
}

/* C++-only helpers the reference exposes to its tests/bench (cpp/src/wholememory/env_func_ptrs.hpp):
 * a cudaMalloc-backed env and a size-class caching env. */
namespace wholememory {
wholememory_env_func_t* get_default_env_func();
wholememory_env_func_t* get_cached_env_func();
void drop_cached_env_func_cache();
}  // namespace wholememory
#endif
