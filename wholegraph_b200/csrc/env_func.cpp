/*
 * Allocation-callback plumbing.
 *  - temp_buffer / output_alloc: how ops call the caller-supplied env functions
 *    (reference cpp/src/wholememory_ops/temp_memory_handle.hpp:23-94, output_memory_handle.hpp:22-94)
 *  - default env (cudaMalloc / cudaMallocHost / malloc) and a cached env (size-class free lists)
 *    for C++ callers (reference cpp/src/wholememory/env_func_ptrs.cpp:90-105, :356-373)
 */
#include "wm_internal.hpp"

#include <unordered_map>

namespace wm {

void* temp_buffer::alloc(size_t elems, wholememory_dtype_t dtype, wholememory_memory_allocation_type_t kind)
{
  release();
  WM_EXPECT(env_ != nullptr, WHOLEMEMORY_INVALID_INPUT, "op needs temporary memory but no env functions were given");
  wholememory_tensor_description_t d;
  wholememory_initialize_tensor_desc(&d);
  d.dim      = 1;
  d.sizes[0] = (int64_t)elems;
  d.dtype    = dtype;
  auto& f    = env_->temporary_fns;
  f.create_memory_context_fn(&ctx_, f.global_context);
  ptr_ = f.malloc_fn(&d, kind, ctx_, f.global_context);
  WM_EXPECT(ptr_ != nullptr || elems == 0, WHOLEMEMORY_OUT_OF_MEMORY, "temporary allocation of %zu elements failed", elems);
  return ptr_;
}

void temp_buffer::release()
{
  if (ctx_ == nullptr) return;
  auto& f = env_->temporary_fns;
  f.free_fn(ctx_, f.global_context);
  f.destroy_memory_context_fn(ctx_, f.global_context);
  ctx_ = nullptr;
  ptr_ = nullptr;
}

void* output_alloc(wholememory_env_func_t* env,
                   void* memory_context,
                   size_t elems,
                   wholememory_dtype_t dtype,
                   wholememory_memory_allocation_type_t kind)
{
  WM_EXPECT(env != nullptr && memory_context != nullptr, WHOLEMEMORY_INVALID_INPUT, "output context / env functions missing");
  wholememory_tensor_description_t d;
  wholememory_initialize_tensor_desc(&d);
  d.dim      = 1;
  d.sizes[0] = (int64_t)elems;
  d.dtype    = dtype;
  return env->output_fns.malloc_fn(&d, kind, memory_context, env->output_fns.global_context);
}

/* ------------------------------------------------------------------ default env */
namespace {

/* memory context of the built-in envs; C++ callers read their variable-size outputs from it */
struct builtin_context {
  wholememory_tensor_description_t desc;
  void* ptr                                  = nullptr;
  wholememory_memory_allocation_type_t kind  = WHOLEMEMORY_MA_NONE;
  size_t capacity                            = 0; /* cached env only */
};

size_t desc_bytes(wholememory_tensor_description_t* d) { return (size_t)wholememory_get_memory_size_from_tensor(d); }

void* raw_alloc(size_t bytes, wholememory_memory_allocation_type_t kind)
{
  void* p = nullptr;
  if (bytes == 0) return nullptr;
  switch (kind) {
    case WHOLEMEMORY_MA_HOST: p = malloc(bytes); break;
    case WHOLEMEMORY_MA_PINNED:
      if (cudaMallocHost(&p, bytes) != cudaSuccess) p = nullptr;
      break;
    case WHOLEMEMORY_MA_DEVICE:
      if (cudaMalloc(&p, bytes) != cudaSuccess) p = nullptr;
      break;
    default: break;
  }
  if (p == nullptr) {
    (void)cudaGetLastError();
    WM_ERROR("builtin env: allocation of %zu bytes (kind %d) failed", bytes, (int)kind);
  }
  return p;
}

void raw_free(void* p, wholememory_memory_allocation_type_t kind)
{
  if (p == nullptr) return;
  switch (kind) {
    case WHOLEMEMORY_MA_HOST: free(p); break;
    case WHOLEMEMORY_MA_PINNED: (void)cudaFreeHost(p); break;
    case WHOLEMEMORY_MA_DEVICE: (void)cudaFree(p); break;
    default: break;
  }
}

void ctx_create(void** ctx, void*)
{
  auto* c = new builtin_context();
  wholememory_initialize_tensor_desc(&c->desc);
  *ctx = c;
}
void ctx_destroy(void* ctx, void*) { delete static_cast<builtin_context*>(ctx); }

void* default_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t kind, void* ctx, void*)
{
  auto* c = static_cast<builtin_context*>(ctx);
  c->desc = *d;
  c->kind = kind;
  c->ptr  = raw_alloc(desc_bytes(d), kind);
  return c->ptr;
}
void default_free(void* ctx, void*)
{
  auto* c = static_cast<builtin_context*>(ctx);
  raw_free(c->ptr, c->kind);
  c->ptr  = nullptr;
  c->kind = WHOLEMEMORY_MA_NONE;
}

wholememory_env_func_t g_default_env = {
  {ctx_create, ctx_destroy, default_malloc, default_free, nullptr},
  {default_malloc, default_free, nullptr},
};

/* ------------------------------------------------------------------ cached env
 * Power-of-two size classes with per-kind free lists: a steady-state training loop performs no
 * cudaMalloc/cudaFree (both synchronise the device) after the first few steps. */
struct block_cache {
  std::mutex mu;
  std::unordered_map<size_t, std::vector<void*>> free_lists[4]; /* by allocation kind */
  void* take(size_t cls, wholememory_memory_allocation_type_t kind)
  {
    std::lock_guard<std::mutex> lk(mu);
    auto& v = free_lists[kind][cls];
    if (v.empty()) return nullptr;
    void* p = v.back();
    v.pop_back();
    return p;
  }
  void give(void* p, size_t cls, wholememory_memory_allocation_type_t kind)
  {
    std::lock_guard<std::mutex> lk(mu);
    free_lists[kind][cls].push_back(p);
  }
  void drop()
  {
    std::lock_guard<std::mutex> lk(mu);
    for (int k = 1; k < 4; ++k) {
      for (auto& kv : free_lists[k])
        for (void* p : kv.second) raw_free(p, (wholememory_memory_allocation_type_t)k);
      free_lists[k].clear();
    }
  }
} g_cache;

size_t size_class(size_t bytes)
{
  size_t c = 256;
  while (c < bytes) c <<= 1;
  return c;
}

void* cached_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t kind, void* ctx, void*)
{
  auto* c     = static_cast<builtin_context*>(ctx);
  c->desc     = *d;
  c->kind     = kind;
  size_t need = desc_bytes(d);
  if (need == 0) {
    c->ptr = nullptr;
    return nullptr;
  }
  c->capacity = size_class(need);
  c->ptr      = g_cache.take(c->capacity, kind);
  if (c->ptr == nullptr) c->ptr = raw_alloc(c->capacity, kind);
  return c->ptr;
}
void cached_free(void* ctx, void*)
{
  auto* c = static_cast<builtin_context*>(ctx);
  if (c->ptr != nullptr) g_cache.give(c->ptr, c->capacity, c->kind);
  c->ptr  = nullptr;
  c->kind = WHOLEMEMORY_MA_NONE;
}

wholememory_env_func_t g_cached_env = {
  {ctx_create, ctx_destroy, cached_malloc, cached_free, nullptr},
  {default_malloc, default_free, nullptr}, /* outputs are owned by the caller: never pooled */
};

}  // namespace
}  // namespace wm

namespace wholememory {
wholememory_env_func_t* get_default_env_func() { return &wm::g_default_env; }
wholememory_env_func_t* get_cached_env_func() { return &wm::g_cached_env; }
void drop_cached_env_func_cache() { wm::g_cache.drop(); }
}  // namespace wholememory

/* Plain-C access to the built-in envs for non-C++ hosts (ctypes / the bench harness).  These four
 * are additions of this build, not reference symbols. */
extern "C" {
wholememory_env_func_t* wgb200_default_env_func() { return &wm::g_default_env; }
wholememory_env_func_t* wgb200_cached_env_func() { return &wm::g_cached_env; }
void wgb200_drop_cached_env_func_cache() { wm::g_cache.drop(); }
}
