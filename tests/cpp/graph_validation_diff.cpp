/*
 * Differential test of the argument checks of the sampling and graph-op entry points --
 *   wholegraph_csr_unweighted_sample_without_replacement, wholegraph_csr_weighted_sample_without_replacement,
 *   graph_append_unique, csr_add_self_loop --
 * the reference's own host files (compiled for the CPU into oracle/_ref/ref_host_graph.so, GPU dispatch targets replaced by
 * stubs returning the sentinel 1000) against this repo's libwholegraph.so.  Same rule as ops_validation_diff.cpp, on a
 * machine WITHOUT a GPU:   reference == 1000  <=>  ours == WHOLEMEMORY_CUDA_ERROR,  otherwise equal error codes.
 * Operands are pointer tensors whose live descriptions are corrupted in one field (rank, last stride, dtype) after creation.
 *
 *   graph_validation_diff <ours.so> <ref_host_graph.so> [iterations]     exit code = number of divergences (capped)
 */
#include <wholememory/env_func_ptrs.h>
#include <wholememory/graph_op.h>
#include <wholememory/tensor_description.h>
#include <wholememory/wholegraph_op.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

namespace {

struct api {
  void* so;
  decltype(&wholememory_make_tensor_from_pointer) from_ptr;
  decltype(&wholememory_destroy_tensor) destroy;
  decltype(&wholememory_tensor_get_tensor_description) get_desc;
  decltype(&wholegraph_csr_unweighted_sample_without_replacement) unweighted;
  decltype(&wholegraph_csr_weighted_sample_without_replacement) weighted;
  decltype(&graph_append_unique) append_unique;
  decltype(&csr_add_self_loop) add_self_loop;
};
template <typename F>
bool sym(void* so, const char* name, F* out)
{
  *out = reinterpret_cast<F>(dlsym(so, name));
  if (*out == nullptr) fprintf(stderr, "missing %s: %s\n", name, dlerror());
  return *out != nullptr;
}
bool load(const char* path, api* a)
{
  a->so = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!a->so) {
    fprintf(stderr, "dlopen(%s): %s\n", path, dlerror());
    return false;
  }
  return sym(a->so, "wholememory_make_tensor_from_pointer", &a->from_ptr) && sym(a->so, "wholememory_destroy_tensor", &a->destroy) &&
         sym(a->so, "wholememory_tensor_get_tensor_description", &a->get_desc) &&
         sym(a->so, "wholegraph_csr_unweighted_sample_without_replacement", &a->unweighted) &&
         sym(a->so, "wholegraph_csr_weighted_sample_without_replacement", &a->weighted) && sym(a->so, "graph_append_unique", &a->append_unique) &&
         sym(a->so, "csr_add_self_loop", &a->add_self_loop);
}

wholememory_tensor_description_t array_desc(int64_t n, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < WHOLEMEMORY_MAX_TENSOR_DIM; ++i) d.sizes[i] = d.strides[i] = 1;
  d.dim = 1, d.dtype = dt, d.sizes[0] = n;
  return d;
}

struct operand {
  wholememory_dtype_t dtype;
  int64_t n;
  int corrupt; /* 0 none, 1 rank 2, 2 rank 0, 3 last stride 2, 4 dtype unknown */
};

wholememory_tensor_t make(api& a, const operand& o, void* mem)
{
  wholememory_tensor_t t = nullptr;
  auto d                 = array_desc(o.n, o.dtype);
  if (a.from_ptr(&t, mem, &d) != WHOLEMEMORY_SUCCESS) return nullptr;
  wholememory_tensor_description_t* live = a.get_desc(t);
  switch (o.corrupt) {
    case 1: live->dim = 2; break;
    case 2: live->dim = 0; break;
    case 3: live->strides[0] = 2; break;
    case 4: live->dtype = WHOLEMEMORY_DT_UNKNOWN; break;
    default: break;
  }
  return t;
}
std::string show(const operand& o) { return "(dtype " + std::to_string((int)o.dtype) + " n " + std::to_string(o.n) + " corrupt " + std::to_string(o.corrupt) + ")"; }

}  // namespace

int main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s <ours.so> <ref_host_graph.so> [iterations]\n", argv[0]);
    return 2;
  }
  api ours{}, ref{};
  if (!load(argv[1], &ours) || !load(argv[2], &ref)) return 2;
  decltype(&fork_get_device_count) devcount;
  decltype(&wholememory_init) init_fn;
  if (!sym(ours.so, "fork_get_device_count", &devcount) || !sym(ours.so, "wholememory_init", &init_fn)) return 2;
  if (devcount() != 0) {
    printf("graph_validation_diff: a GPU is present, nothing compared (the equivalence holds on a GPU-less machine only)\n");
    return 0;
  }
  init_fn(0, LEVEL_FATAL);
  const long iters = argc > 3 ? atol(argv[3]) : 100000;
  std::mt19937_64 rng(4242);
  auto pick = [&](long lo, long hi) { return lo + (long)(rng() % (uint64_t)(hi - lo + 1)); };
  static char mem[5][1 << 12];
  static char fake_ctx[64];
  const wholememory_dtype_t dts[] = {WHOLEMEMORY_DT_INT, WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_DOUBLE, WHOLEMEMORY_DT_HALF};
  const char* names[4] = {"unweighted_sample", "weighted_sample", "append_unique", "csr_add_self_loop"};
  int divergences = 0;
  long dispatched = 0, refused = 0;

  for (long it = 0; it < iters; ++it) {
    operand ops[5];
    for (auto& o : ops) o = {dts[pick(0, 4)], pick(0, 30), pick(0, 2) == 0 ? (int)pick(1, 4) : 0};
    /* make "plausible" calls common: row_ptr int64, offsets int32 ... two thirds of the time */
    if (pick(0, 2) != 0) ops[0].dtype = WHOLEMEMORY_DT_INT64, ops[1].dtype = pick(0, 1) ? WHOLEMEMORY_DT_INT : WHOLEMEMORY_DT_INT64,
                         ops[2].dtype = WHOLEMEMORY_DT_FLOAT, ops[3].dtype = ops[1].dtype, ops[4].dtype = WHOLEMEMORY_DT_INT;
    const bool none_mapping = pick(0, 3) == 0; /* graph_append_unique: the "None" mapping is a 0-dim tensor */
    int codes[2][4];
    for (int side = 0; side < 2; ++side) {
      api& a = side == 0 ? ours : ref;
      wholememory_tensor_t t[5];
      for (int k = 0; k < 5; ++k) {
        t[k] = make(a, ops[k], mem[k]);
        if (t[k] == nullptr) {
          fprintf(stderr, "setup failed\n");
          return 2;
        }
      }
      /* roles: 0 row_ptr, 1 col, 2 weights, 3 centers, 4 offsets */
      codes[side][0] = (int)a.unweighted(t[0], t[1], t[3], 10, t[4], fake_ctx, nullptr, nullptr, 1ull, nullptr, nullptr);
      codes[side][1] = (int)a.weighted(t[0], t[1], t[2], t[3], 10, t[4], fake_ctx, nullptr, nullptr, 1ull, nullptr, nullptr);
      /* roles: 1 targets, 3 neighbors, 4 mapping (or a 0-dim "None") */
      wholememory_tensor_description_t* md = a.get_desc(t[4]);
      const int saved_dim                  = md->dim;
      if (none_mapping) md->dim = 0;
      codes[side][2] = (int)a.append_unique(t[1], t[3], fake_ctx, t[4], nullptr, nullptr);
      md->dim        = saved_dim;
      /* roles: 0,1 input CSR, 3,4 output CSR */
      codes[side][3] = (int)a.add_self_loop(t[0], t[1], t[3], t[4], nullptr);
      for (auto x : t) a.destroy(x);
    }
    for (int op = 0; op < 4; ++op) {
      const int mine = codes[0][op], theirs = codes[1][op];
      const bool same = theirs == 1000 ? mine == (int)WHOLEMEMORY_CUDA_ERROR : mine == theirs;
      (theirs == 1000 ? dispatched : refused)++;
      if (!same && ++divergences <= 40)
        fprintf(stderr, "DIVERGENCE [%s] iteration %ld: ours %d reference %d | %s %s %s %s %s none_mapping %d\n", names[op], it, mine, theirs,
                show(ops[0]).c_str(), show(ops[1]).c_str(), show(ops[2]).c_str(), show(ops[3]).c_str(), show(ops[4]).c_str(), (int)none_mapping);
    }
  }
  printf("graph_validation_diff: %ld iterations, %d divergences (%ld calls dispatched by the reference, %ld refused)\n", iters, divergences,
         dispatched, refused);
  return divergences > 100 ? 100 : divergences;
}
