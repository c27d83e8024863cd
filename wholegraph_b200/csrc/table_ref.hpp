/* Plain (host + device) structs handed to the row-move kernels by value. */
#pragma once
#include <stdint.h>

namespace wm {

constexpr int kMaxInlineRanks = 16;

/* How a kernel finds table row bytes.  Passed BY VALUE as a kernel parameter. */
struct table_ref {
  enum : int { FLAT = 0, CHUNK_REGULAR = 1, CHUNK_IRREGULAR = 2, DEVTAB_REGULAR = 3, DEVTAB_IRREGULAR = 4 };
  int mode;
  int nranks;
  int has_remote; /* some rows live outside this GPU's HBM (peer GPUs over NVLink, or host memory) */
  uint64_t chunk_bytes;                        /* *_REGULAR: bytes owned by each rank */
  char* base[kMaxInlineRanks];                 /* FLAT: base[0]; CHUNK_*: start of rank r's partition */
  uint64_t first_byte[kMaxInlineRanks + 1];    /* CHUNK_IRREGULAR: partition start offsets */
  char* const* dev_bases;                      /* DEVTAB_*: public gref tables in device memory */
  const uint64_t* dev_first_byte;
};

/* strided matrix geometry in BYTES (host computes these once) */
struct row_geom {
  int64_t table_offset_bytes; /* storage_offset * esize */
  int64_t table_stride_bytes;
  int64_t dense_stride_bytes;
  int row_elems;   /* columns */
  int batch_rows;  /* rows resolved per warp batch: power of two <= 32 */
  /* division-free "unit index -> (row, unit in row)": row = (w * div_magic) >> 40, exact for w, units < 2^20.
   * (an integer divide per vector runs on the XU pipe and was measured to saturate it at 96 %) */
  uint64_t div_magic;
  int units_per_row; /* vectors (copy kernels) or ALIGN-packs (converting kernels) per row */
  int policy;        /* load variant | store variant << 4; 0 = streaming loads (local rows), 2 = L1-allocating loads (rows that can be remote) */
};

}  // namespace wm
