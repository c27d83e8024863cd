/* Shared launcher template for the converting row-move kernels (instantiated per dtype family). */
#pragma once
#include "gather_scatter.cuh"
#include "ops_internal.hpp"

namespace wm {

template <typename TableT, typename DenseT, typename IdxT, bool GATHER>
void launch_cvt_align(int align, const table_ref& t, const row_geom& g, const void* idx, int64_t n, char* dense, int grid, cudaStream_t s)
{
  constexpr int kMaxAlign = 16 / (sizeof(TableT) > sizeof(DenseT) ? sizeof(TableT) : sizeof(DenseT));
  const IdxT* ip          = static_cast<const IdxT*>(idx);
  if constexpr (kMaxAlign >= 8) {
    if (align >= 8) {
      row_move_cvt_kernel<TableT, DenseT, IdxT, 8, GATHER><<<grid, 256, 0, s>>>(t, g, ip, n, dense);
      return;
    }
  }
  if constexpr (kMaxAlign >= 4) {
    if (align >= 4) {
      row_move_cvt_kernel<TableT, DenseT, IdxT, 4, GATHER><<<grid, 256, 0, s>>>(t, g, ip, n, dense);
      return;
    }
  }
  if constexpr (kMaxAlign >= 2) {
    if (align >= 2) {
      row_move_cvt_kernel<TableT, DenseT, IdxT, 2, GATHER><<<grid, 256, 0, s>>>(t, g, ip, n, dense);
      return;
    }
  }
  row_move_cvt_kernel<TableT, DenseT, IdxT, 1, GATHER><<<grid, 256, 0, s>>>(t, g, ip, n, dense);
}

template <typename TableT, typename DenseT>
void launch_cvt(bool gather, const table_ref& t, const row_geom& g, const void* idx, bool idx64, int64_t n, char* dense, int align, int grid, cudaStream_t s)
{
  if (gather) {
    if (idx64) launch_cvt_align<TableT, DenseT, int64_t, true>(align, t, g, idx, n, dense, grid, s);
    else launch_cvt_align<TableT, DenseT, int32_t, true>(align, t, g, idx, n, dense, grid, s);
  } else {
    if (idx64) launch_cvt_align<TableT, DenseT, int64_t, false>(align, t, g, idx, n, dense, grid, s);
    else launch_cvt_align<TableT, DenseT, int32_t, false>(align, t, g, idx, n, dense, grid, s);
  }
}

}  // namespace wm
