/*
 * wholememory_env_test_op: allocator-plumbing self test used by the binding's unit test.
 * Same observable behaviour as reference cpp/src/wholememory_ops/wholememory_test_op.cu:24-165:
 * out[i, :] = T(float(i)) + input[:], written to the fixed output and to three variable outputs
 * allocated through the output callbacks (device, pinned, host).
 */
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "wm_internal.hpp"

#include <algorithm>

namespace wm {
namespace {

template <typename T>
__global__ void add_row_id_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t dim, int64_t out_stride)
{
  T* row       = out + out_stride * blockIdx.x;
  const T base = static_cast<T>(static_cast<float>(blockIdx.x));
  for (int64_t c = threadIdx.x; c < dim; c += blockDim.x) row[c] = base + in[c];
}

template <typename T>
void run(const void* in, void* out, int64_t dim, int64_t rows, int64_t stride, cudaStream_t s)
{
  if (rows == 0 || dim == 0) return;
  add_row_id_kernel<T><<<(unsigned)rows, (unsigned)std::min<int64_t>(dim, 512), 0, s>>>(static_cast<const T*>(in), static_cast<T*>(out), dim, stride);
}

}  // namespace
}  // namespace wm

extern "C" wholememory_error_code_t wholememory_env_test_op(wholememory_tensor_t input_tensor,
                                                            wholememory_tensor_t output_fixed_tensor,
                                                            void* output_variable_device_tensor_handle,
                                                            void* output_variable_pinned_tensor_handle,
                                                            void* output_variable_host_tensor_handle,
                                                            int64_t output_variable_entry_count,
                                                            wholememory_env_func_t* p_env_fns,
                                                            void* stream)
{
  return wm::guarded("wholememory_env_test_op", [&]() -> wholememory_error_code_t {
    using namespace wm;
    WM_REQUIRE_LIVE(input_tensor);
    WM_REQUIRE_LIVE(output_fixed_tensor);
    require_cuda("wholememory_env_test_op");
    auto* id = wholememory_tensor_get_tensor_description(input_tensor);
    auto* od = wholememory_tensor_get_tensor_description(output_fixed_tensor);
    WM_EXPECT(id->dim == 1 && od->dim == 2 && od->sizes[0] == output_variable_entry_count && od->sizes[1] == id->sizes[0] &&
                id->dtype == od->dtype,
              WHOLEMEMORY_INVALID_INPUT, "env_test_op: shape / dtype mismatch");
    const int64_t dim = id->sizes[0], rows = output_variable_entry_count;
    const size_t es   = wholememory_dtype_get_element_size(id->dtype);
    auto s            = static_cast<cudaStream_t>(stream);
    temp_buffer tmp(p_env_fns);
    void* t        = tmp.device((size_t)(rows * dim), id->dtype);
    const void* in = wholememory_tensor_get_data_pointer(input_tensor);
    switch (id->dtype) {
      case WHOLEMEMORY_DT_FLOAT: run<float>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_DOUBLE: run<double>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_HALF: run<__half>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_BF16: run<__nv_bfloat16>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_INT: run<int32_t>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_INT64: run<int64_t>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_INT16: run<int16_t>(in, t, dim, rows, dim, s); break;
      case WHOLEMEMORY_DT_INT8: run<int8_t>(in, t, dim, rows, dim, s); break;
      default: return WHOLEMEMORY_INVALID_INPUT;
    }
    WM_CUDA(cudaGetLastError());
    const size_t row_bytes = (size_t)dim * es;
    /* fixed output honours its row stride */
    WM_CUDA(cudaMemcpy2DAsync(wholememory_tensor_get_data_pointer(output_fixed_tensor), (size_t)od->strides[0] * es, t, row_bytes,
                              row_bytes, (size_t)rows, cudaMemcpyDeviceToDevice, s));
    struct {
      void* ctx;
      wholememory_memory_allocation_type_t kind;
      cudaMemcpyKind copy;
    } outs[3] = {{output_variable_device_tensor_handle, WHOLEMEMORY_MA_DEVICE, cudaMemcpyDeviceToDevice},
                 {output_variable_pinned_tensor_handle, WHOLEMEMORY_MA_PINNED, cudaMemcpyDeviceToHost},
                 {output_variable_host_tensor_handle, WHOLEMEMORY_MA_HOST, cudaMemcpyDeviceToHost}};
    for (auto& o : outs) {
      if (o.ctx == nullptr) continue;
      wholememory_tensor_description_t d = *od;
      d.strides[0]                       = dim;
      d.storage_offset                   = 0;
      void* p = p_env_fns->output_fns.malloc_fn(&d, o.kind, o.ctx, p_env_fns->output_fns.global_context);
      WM_EXPECT(p != nullptr || rows * dim == 0, WHOLEMEMORY_OUT_OF_MEMORY, "env_test_op: output allocation failed");
      WM_CUDA(cudaMemcpyAsync(p, t, row_bytes * (size_t)rows, o.copy, s));
    }
    WM_CUDA(cudaStreamSynchronize(s));
    return WHOLEMEMORY_SUCCESS;
  });
}
