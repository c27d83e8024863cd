/*
 * B200-native WholeMemory engine -- C ABI boundary (drop-in for libwholegraph).
 *
 * wholememory_gref_t: the by-value "global reference" a kernel uses to turn a global element
 * index into an address.  Layout (40 bytes, field order, meaning) is ABI and matches
 * reference cpp/include/wholememory/global_reference.h:32-42 because downstream kernels
 * (cugraph-ops style callers using device_reference.cuh) receive it by value.
 */
#pragma once
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct wholememory_gref_t {
  void* pointer;               /* stride==0: flat base VA.  stride>0: DEVICE array of world_size chunk bases */
  size_t* rank_memory_offsets; /* DEVICE array, world_size+1 byte offsets (chunked, !same_chunk lookup) */
  int world_size;
  size_t stride;   /* 0 => flat; else bytes owned by each rank when same_chunk */
  bool same_chunk; /* true: owner = byte_offset / stride */
};

/* replaces reference global_reference.h:49 / global_reference.cpp */
wholememory_gref_t wholememory_create_continuous_global_reference(void* ptr);

/* ABI placeholder only: the NVSHMEM backend is out of scope (reference global_reference.h:51-58). */
struct wholememory_nvshmem_ref_t {
  void* pointer;
  size_t* rank_memory_offsets;
  size_t stride;
  int world_rank;
  int world_size;
  bool same_chunk;
};

#ifdef __cplusplus
}
#endif
