/*
 * WholeMemory C ABI: init, communicators, memory handles.
 * Drop-in boundary for reference cpp/include/wholememory/wholememory.h (enum values,
 * struct layouts, symbol names and argument order are ABI; citations per entry point).
 * Implementation: wholegraph_b200/csrc/{runtime,communicator,memory_handle}.cpp.
 *
 * Scope of this build: ONE NVSwitch box, one process per GPU.  Multi-node,
 * MNNVL cliques, HIERARCHY memory and NVSHMEM are out of scope; their entry points exist and
 * return WHOLEMEMORY_NOT_IMPLEMENTED / NOT_SUPPORTED.
 */
#pragma once
#include <stdio.h>
#include <unistd.h>
#include <wholememory/global_reference.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference wholememory.h:32-44 */
enum wholememory_error_code_t {
  WHOLEMEMORY_SUCCESS = 0,
  WHOLEMEMORY_UNKNOW_ERROR,
  WHOLEMEMORY_NOT_IMPLEMENTED,
  WHOLEMEMORY_LOGIC_ERROR,
  WHOLEMEMORY_CUDA_ERROR,
  WHOLEMEMORY_COMMUNICATION_ERROR,
  WHOLEMEMORY_INVALID_INPUT,
  WHOLEMEMORY_INVALID_VALUE,
  WHOLEMEMORY_OUT_OF_MEMORY,
  WHOLEMEMORY_NOT_SUPPORTED,
  WHOLEMEMORY_SYSTEM_ERROR,
};

/* early-return helper used by C++ callers of the ABI (reference wholememory.h:46-54) */
#define WHOLEMEMORY_RETURN_ON_FAIL(X)                                                          \
  do {                                                                                         \
    auto wm_rc_ = (X);                                                                         \
    if (wm_rc_ != WHOLEMEMORY_SUCCESS) {                                                       \
      fprintf(stderr, "File %s line %d %s failed.\n", __FILE__, __LINE__, #X);                 \
      return wm_rc_;                                                                           \
    }                                                                                          \
  } while (0)

/* reference wholememory.h:59-65 */
enum wholememory_memory_type_t {
  WHOLEMEMORY_MT_NONE = 0,
  WHOLEMEMORY_MT_CONTINUOUS,  /* every rank's shard mapped into one flat VA range on every GPU */
  WHOLEMEMORY_MT_CHUNKED,     /* every rank's shard mapped, one VA range per shard */
  WHOLEMEMORY_MT_DISTRIBUTED, /* shards not visible through the ABI; ops exchange rows */
  WHOLEMEMORY_MT_HIERARCHY,   /* multi-node two-level; not supported by this build */
};

/* reference wholememory.h:70-74 */
enum wholememory_memory_location_t {
  WHOLEMEMORY_ML_NONE = 0,
  WHOLEMEMORY_ML_DEVICE,
  WHOLEMEMORY_ML_HOST,
};

/* reference wholememory.h:76-80 */
enum wholememory_distributed_backend_t {
  WHOLEMEMORY_DB_NONE = 0,
  WHOLEMEMORY_DB_NCCL,
  WHOLEMEMORY_DB_NVSHMEM, /* not supported by this build */
};

/* reference wholememory.h:82-89 */
enum LogLevel { LEVEL_FATAL = 0, LEVEL_ERROR, LEVEL_WARN, LEVEL_INFO, LEVEL_DEBUG, LEVEL_TRACE };

#define WHOLEMEMORY_SPILT_NO_COLOR -1

/* reference wholememory.h:97-109.  init: flags must be 0. */
wholememory_error_code_t wholememory_init(unsigned int flags, LogLevel log_level = LEVEL_INFO);
wholememory_error_code_t wholememory_finalize();

typedef struct wholememory_comm_* wholememory_comm_t;

/* reference wholememory.h:117-124 (MNNVL clique description; this build reports "not in a clique") */
struct clique_info_t {
  int is_in_clique;
  int clique_first_rank;
  int clique_rank;
  int clique_rank_num;
  int clique_id;
  int clique_num;
};

#define WHOLEMEMORY_UNIQUE_ID_BYTES (128)
struct wholememory_unique_id_t {
  char internal[WHOLEMEMORY_UNIQUE_ID_BYTES];
};

/* ---- communicators (reference wholememory.h:137-245) ---- */
wholememory_error_code_t wholememory_create_unique_id(wholememory_unique_id_t* unique_id);
/* collective over `size` processes that all pass the same unique_id */
wholememory_error_code_t wholememory_create_communicator(wholememory_comm_t* comm, wholememory_unique_id_t unique_id,
    int rank, int size);
wholememory_error_code_t wholememory_split_communicator(wholememory_comm_t* new_comm, wholememory_comm_t comm,
    int color, int key);
/* also frees every WholeMemory handle still allocated on it */
wholememory_error_code_t wholememory_destroy_communicator(wholememory_comm_t comm);
/* WHOLEMEMORY_SUCCESS when the (type, location) pair can be allocated on comm */
wholememory_error_code_t wholememory_communicator_support_type_location(wholememory_comm_t comm,
    wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location);
wholememory_error_code_t wholememory_communicator_get_rank(int* rank, wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_get_size(int* size, wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_get_local_size(int* local_size, wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_get_clique_info(clique_info_t* clique_info,
    wholememory_comm_t comm);
bool wholememory_communicator_is_bind_to_nvshmem(wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_set_distributed_backend(wholememory_comm_t comm,
    wholememory_distributed_backend_t distributed_backend);
wholememory_distributed_backend_t wholememory_communicator_get_distributed_backend(wholememory_comm_t comm);
wholememory_error_code_t wholememory_communicator_barrier(wholememory_comm_t comm);
bool wholememory_is_intranode_communicator(wholememory_comm_t comm);
bool wholememory_is_intra_mnnvl_communicator(wholememory_comm_t comm);
bool wholememory_is_build_with_nvshmem();

/* ---- memory handles (reference wholememory.h:247-440) ---- */
typedef struct wholememory_handle_* wholememory_handle_t;

/* Collective.  total_size and data_granularity in bytes; rank_entry_partition (optional) gives
 * the number of data_granularity-sized entries owned by each rank. */
wholememory_error_code_t wholememory_malloc(wholememory_handle_t* handle_out, size_t total_size,
    wholememory_comm_t comm, wholememory_memory_type_t memory_type, wholememory_memory_location_t memory_location,
    size_t data_granularity, size_t* rank_entry_partition = nullptr);
wholememory_error_code_t wholememory_free(wholememory_handle_t handle); /* collective */

wholememory_error_code_t wholememory_get_communicator(wholememory_comm_t* comm, wholememory_handle_t handle);
/* HIERARCHY only => WHOLEMEMORY_NOT_SUPPORTED here */
wholememory_error_code_t wholememory_get_local_communicator(wholememory_comm_t* comm, wholememory_handle_t handle);
wholememory_error_code_t wholememory_get_cross_communicator(wholememory_comm_t* comm, wholememory_handle_t handle);

wholememory_memory_type_t wholememory_get_memory_type(wholememory_handle_t handle);
wholememory_memory_location_t wholememory_get_memory_location(wholememory_handle_t handle);
wholememory_distributed_backend_t wholememory_get_distributed_backend(wholememory_handle_t handle);
size_t wholememory_get_total_size(wholememory_handle_t handle);
size_t wholememory_get_data_granularity(wholememory_handle_t handle);

/* the shard this rank is responsible for: pointer, bytes, byte offset in the whole memory */
wholememory_error_code_t wholememory_get_local_memory(void** local_ptr, size_t* local_size, size_t* local_offset,
    wholememory_handle_t handle);
wholememory_error_code_t wholememory_get_local_size(size_t* local_size, wholememory_handle_t handle);
wholememory_error_code_t wholememory_get_local_offset(size_t* local_offset, wholememory_handle_t handle);
/* mapped types only: where rank's shard is visible in THIS process */
wholememory_error_code_t wholememory_get_rank_memory(void** rank_memory_ptr, size_t* rank_memory_size,
    size_t* rank_memory_offset, int rank, wholememory_handle_t handle);
/* entries per rank of the default partition = ceil(total / world_size) */
wholememory_error_code_t wholememory_equal_entry_partition_plan(size_t* entry_per_rank, size_t total_entry_count,
    int world_size);
/* CONTINUOUS only */
wholememory_error_code_t wholememory_get_global_pointer(void** global_ptr, wholememory_handle_t handle);
/* CONTINUOUS or CHUNKED */
wholememory_error_code_t wholememory_get_global_reference(wholememory_gref_t* gref, wholememory_handle_t handle);
/* world_size sizes / world_size+1 offsets, in bytes */
wholememory_error_code_t wholememory_get_rank_partition_sizes(size_t* rank_mem_sizes, wholememory_handle_t handle);
wholememory_error_code_t wholememory_get_rank_partition_offsets(size_t* rank_mem_offsets,
    wholememory_handle_t handle);

/* GPU count probed in a forked child so the caller never creates a CUDA context */
int fork_get_device_count();

/* part-file load / store (reference :448-470; wholegraph_b200/csrc/file_io.cpp).  round_robin_size != 0 is refused. */
wholememory_error_code_t wholememory_load_from_file(wholememory_handle_t handle, size_t memory_offset,
    size_t memory_entry_size, size_t file_entry_size, const char** file_names, int file_count, int round_robin_size);
wholememory_error_code_t wholememory_store_to_file(wholememory_handle_t handle, size_t memory_offset,
    size_t memory_entry_stride, size_t file_entry_size, const char* local_file_name);

#ifdef __cplusplus
}
#endif
