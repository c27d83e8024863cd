/*
 * Native torch allocator callbacks ("env functions") for libwholegraph.so.
 *
 * Every temporary and every variable-size output an op needs is a torch tensor obtained from
 * torch's caching allocator.  The Python-callback version of these four functions
 * (wholegraph_b200/torch/wholegraph_env.py) re-enters the interpreter once per allocation; this
 * module provides the same protocol as plain C++ functions, so an op call crosses into Python
 * zero times.  Functionally replaces the reference's optional extension
 * python/pylibwholegraph/pylibwholegraph/torch_cpp_ext/ (torch_env_func_ptrs.cpp:25-49 env table,
 * torch_utils.cpp malloc/free, wholegraph_torch_ext.cpp:50-66 module surface) -- same six Python
 * entry points, own implementation.  It depends only on the ABI headers (struct layouts), not on
 * libwholegraph.so itself.
 *
 * Built in-tree by wholegraph_b200/csrc/torch_ext/build.sh into wholegraph_b200/lib/.
 */
#include <torch/extension.h>

#include <c10/cuda/CUDAStream.h>

#include <wholememory/env_func_ptrs.h>
#include <wholememory/tensor_description.h>

#include <atomic>
#include <vector>

namespace {

/* One allocation made on behalf of the library.  Temporaries: created and destroyed by the library
 * through the callbacks.  Outputs: created by the Python caller (create_output_context), filled by
 * the library, read back with get_tensor_from_context. */
struct tensor_context {
  at::Tensor tensor;
};

std::atomic<int64_t> g_live_contexts{0};

bool to_scalar_type(wholememory_dtype_t dt, at::ScalarType* out)
{
  switch (dt) {
    case WHOLEMEMORY_DT_FLOAT: *out = at::kFloat; return true;
    case WHOLEMEMORY_DT_HALF: *out = at::kHalf; return true;
    case WHOLEMEMORY_DT_DOUBLE: *out = at::kDouble; return true;
    case WHOLEMEMORY_DT_BF16: *out = at::kBFloat16; return true;
    case WHOLEMEMORY_DT_INT: *out = at::kInt; return true;
    case WHOLEMEMORY_DT_INT64: *out = at::kLong; return true;
    case WHOLEMEMORY_DT_INT16: *out = at::kShort; return true;
    case WHOLEMEMORY_DT_INT8: *out = at::kChar; return true;
    default: return false;
  }
}

void create_context_fn(void** memory_context, void* /*global_context*/)
{
  *memory_context = new tensor_context();
  g_live_contexts.fetch_add(1, std::memory_order_relaxed);
}

void destroy_context_fn(void* memory_context, void* /*global_context*/)
{
  if (memory_context == nullptr) return;
  delete static_cast<tensor_context*>(memory_context);
  g_live_contexts.fetch_sub(1, std::memory_order_relaxed);
}

/* C callback: must not throw.  A failed allocation returns nullptr, which the library turns into
 * WHOLEMEMORY_OUT_OF_MEMORY. */
void* malloc_fn(wholememory_tensor_description_t* desc,
                wholememory_memory_allocation_type_t kind,
                void* memory_context,
                void* /*global_context*/)
{
  auto* ctx = static_cast<tensor_context*>(memory_context);
  if (ctx == nullptr || desc == nullptr || desc->dim < 0 || desc->dim > WHOLEMEMORY_MAX_TENSOR_DIM) return nullptr;
  at::ScalarType st;
  if (!to_scalar_type(desc->dtype, &st)) return nullptr;
  std::vector<int64_t> shape(desc->sizes, desc->sizes + desc->dim);
  try {
    auto opts = at::TensorOptions().dtype(st);
    switch (kind) {
      case WHOLEMEMORY_MA_DEVICE: opts = opts.device(at::kCUDA); break; /* current device */
      case WHOLEMEMORY_MA_HOST: opts = opts.device(at::kCPU); break;
      case WHOLEMEMORY_MA_PINNED: opts = opts.device(at::kCPU).pinned_memory(true); break;
      default: return nullptr;
    }
    ctx->tensor = at::empty(shape, opts);
    return ctx->tensor.data_ptr();
  } catch (const std::exception& e) {
    fprintf(stderr, "[wholegraph_b200 torch env] allocation failed: %s\n", e.what());
    ctx->tensor = at::Tensor();
    return nullptr;
  } catch (...) {
    ctx->tensor = at::Tensor();
    return nullptr;
  }
}

void free_fn(void* memory_context, void* /*global_context*/)
{
  if (memory_context == nullptr) return;
  static_cast<tensor_context*>(memory_context)->tensor = at::Tensor();
}

wholememory_env_func_t g_env = {
  {create_context_fn, destroy_context_fn, malloc_fn, free_fn, nullptr},
  {malloc_fn, free_fn, nullptr},
};

/* ---- Python surface (same names as the reference extension) ---- */
int64_t get_wholegraph_env_fns() { return reinterpret_cast<int64_t>(&g_env); }

int64_t get_stream()
{
  return reinterpret_cast<int64_t>(static_cast<void*>(c10::cuda::getCurrentCUDAStream().stream()));
}

int64_t create_output_context()
{
  void* ctx = nullptr;
  create_context_fn(&ctx, nullptr);
  return reinterpret_cast<int64_t>(ctx);
}

void destroy_output_context(int64_t ctx) { destroy_context_fn(reinterpret_cast<void*>(ctx), nullptr); }

void free_context_data(int64_t ctx) { free_fn(reinterpret_cast<void*>(ctx), nullptr); }

py::object get_tensor_from_context(int64_t ctx)
{
  auto* c = reinterpret_cast<tensor_context*>(ctx);
  if (c == nullptr || !c->tensor.defined()) return py::none();
  return py::cast(c->tensor);
}

int64_t live_context_count() { return g_live_contexts.load(std::memory_order_relaxed); }

}  // namespace

PYBIND11_MODULE(wholegraph_b200_torch_ext, m)
{
  m.doc() = "torch-backed allocation callbacks for libwholegraph.so without Python re-entry";
  m.def("get_wholegraph_env_fns", &get_wholegraph_env_fns, "address of the wholememory_env_func_t table");
  m.def("get_stream", &get_stream, "current CUDA stream of the current device as an integer");
  m.def("create_output_context", &create_output_context, "new output memory context (caller-owned)");
  m.def("destroy_output_context", &destroy_output_context, "delete an output memory context");
  m.def("free_context_data", &free_context_data, "drop the tensor held by a context");
  m.def("get_tensor_from_context", &get_tensor_from_context, "tensor the library allocated into the context (None if empty)");
  m.def("live_context_count", &live_context_count, "contexts created and not yet destroyed (leak check)");
}
