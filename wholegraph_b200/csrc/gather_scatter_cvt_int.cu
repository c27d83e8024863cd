/* Converting gather/scatter kernels, integer family {int8, int16, int32, int64}
 * (reference gather_func_impl_integer_data_*.cu registers ALLSINT x ALLSINT). */
#include "gather_scatter_cvt.cuh"

namespace wm {

#define WM_PAIR(TDT, DDT, TT, DT) \
  if (table_dt == TDT && dense_dt == DDT) return &launch_cvt<TT, DT>;

cvt_launch_fn find_int_cvt(wholememory_dtype_t table_dt, wholememory_dtype_t dense_dt)
{
  WM_PAIR(WHOLEMEMORY_DT_INT8, WHOLEMEMORY_DT_INT16, int8_t, int16_t)
  WM_PAIR(WHOLEMEMORY_DT_INT8, WHOLEMEMORY_DT_INT, int8_t, int32_t)
  WM_PAIR(WHOLEMEMORY_DT_INT8, WHOLEMEMORY_DT_INT64, int8_t, int64_t)
  WM_PAIR(WHOLEMEMORY_DT_INT16, WHOLEMEMORY_DT_INT8, int16_t, int8_t)
  WM_PAIR(WHOLEMEMORY_DT_INT16, WHOLEMEMORY_DT_INT, int16_t, int32_t)
  WM_PAIR(WHOLEMEMORY_DT_INT16, WHOLEMEMORY_DT_INT64, int16_t, int64_t)
  WM_PAIR(WHOLEMEMORY_DT_INT, WHOLEMEMORY_DT_INT8, int32_t, int8_t)
  WM_PAIR(WHOLEMEMORY_DT_INT, WHOLEMEMORY_DT_INT16, int32_t, int16_t)
  WM_PAIR(WHOLEMEMORY_DT_INT, WHOLEMEMORY_DT_INT64, int32_t, int64_t)
  WM_PAIR(WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_INT8, int64_t, int8_t)
  WM_PAIR(WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_INT16, int64_t, int16_t)
  WM_PAIR(WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_INT, int64_t, int32_t)
  return nullptr;
}

}  // namespace wm
