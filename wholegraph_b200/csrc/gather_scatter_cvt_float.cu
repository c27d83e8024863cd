/* Converting gather/scatter kernels, floating-point family {fp16, bf16, fp32, fp64}.
 * Registered pairs in the reference: {half,float,double}^2 (gather_func_impl_floating_data_*.cu:49-52);
 * bf16 is added here (the conversion chain through float is the one type_caster<__nv_bfloat16> defines). */
#include "gather_scatter_cvt.cuh"

namespace wm {

int cvt_blocks_per_sm()
{
  static int occ = [] {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, row_move_cvt_kernel<__half, float, int64_t, 4, true>, 256, 0) != cudaSuccess || o <= 0) {
      (void)cudaGetLastError();
      o = 4;
    }
    return o;
  }();
  return occ;
}

#define WM_PAIR(TDT, DDT, TT, DT) \
  if (table_dt == TDT && dense_dt == DDT) return &launch_cvt<TT, DT>;

cvt_launch_fn find_float_cvt(wholememory_dtype_t table_dt, wholememory_dtype_t dense_dt)
{
  WM_PAIR(WHOLEMEMORY_DT_HALF, WHOLEMEMORY_DT_FLOAT, __half, float)
  WM_PAIR(WHOLEMEMORY_DT_HALF, WHOLEMEMORY_DT_DOUBLE, __half, double)
  WM_PAIR(WHOLEMEMORY_DT_HALF, WHOLEMEMORY_DT_BF16, __half, __nv_bfloat16)
  WM_PAIR(WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_HALF, float, __half)
  WM_PAIR(WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_DOUBLE, float, double)
  WM_PAIR(WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_BF16, float, __nv_bfloat16)
  WM_PAIR(WHOLEMEMORY_DT_DOUBLE, WHOLEMEMORY_DT_HALF, double, __half)
  WM_PAIR(WHOLEMEMORY_DT_DOUBLE, WHOLEMEMORY_DT_FLOAT, double, float)
  WM_PAIR(WHOLEMEMORY_DT_DOUBLE, WHOLEMEMORY_DT_BF16, double, __nv_bfloat16)
  WM_PAIR(WHOLEMEMORY_DT_BF16, WHOLEMEMORY_DT_HALF, __nv_bfloat16, __half)
  WM_PAIR(WHOLEMEMORY_DT_BF16, WHOLEMEMORY_DT_FLOAT, __nv_bfloat16, float)
  WM_PAIR(WHOLEMEMORY_DT_BF16, WHOLEMEMORY_DT_DOUBLE, __nv_bfloat16, double)
  return nullptr;
}

}  // namespace wm
