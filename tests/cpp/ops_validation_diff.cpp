/*
 * Differential test of the argument checks of wholememory_gather / wholememory_scatter: the reference's own
 * gather_op.cpp + scatter_op.cpp (compiled for the CPU into oracle/_ref/ref_host_ops.so, with the GPU functions they
 * dispatch to replaced by stubs that return the sentinel 1000 = "checks passed, dispatched") against this repo's
 * libwholegraph.so.  Runs on a machine WITHOUT a GPU: there this repo's library answers WHOLEMEMORY_CUDA_ERROR exactly
 * when its checks pass and the kernel launch would follow ("no CPU fallback"), so
 *        reference == 1000   <=>   ours == WHOLEMEMORY_CUDA_ERROR,   otherwise the two error codes must be equal.
 * Operands are pointer tensors; their descriptions are also corrupted AFTER creation (the accessor hands out a mutable
 * pointer) to reach the checks a well-formed tensor cannot fail.
 *
 *   ops_validation_diff <ours.so> <ref_host_ops.so> [iterations]      exit code = number of divergences (capped)
 */
#include <wholememory/env_func_ptrs.h>
#include <wholememory/tensor_description.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
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
  decltype(&wholememory_gather) gather;
  decltype(&wholememory_scatter) scatter;
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
         sym(a->so, "wholememory_tensor_get_tensor_description", &a->get_desc) && sym(a->so, "wholememory_gather", &a->gather) &&
         sym(a->so, "wholememory_scatter", &a->scatter);
}

wholememory_tensor_description_t make_desc(int dim, int64_t rows, int64_t cols, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < WHOLEMEMORY_MAX_TENSOR_DIM; ++i) d.sizes[i] = d.strides[i] = 1;
  d.dim = dim, d.dtype = dt, d.storage_offset = 0;
  if (dim == 1) {
    d.sizes[0] = rows, d.strides[0] = 1;
  } else {
    d.sizes[0] = rows, d.sizes[1] = cols, d.strides[0] = cols, d.strides[1] = 1;
  }
  return d;
}
std::string show(const wholememory_tensor_description_t& d)
{
  return "dim " + std::to_string(d.dim) + " dtype " + std::to_string((int)d.dtype) + " sizes " + std::to_string(d.sizes[0]) + "," +
         std::to_string(d.sizes[1]) + " strides " + std::to_string(d.strides[0]) + "," + std::to_string(d.strides[1]);
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s <ours.so> <ref_host_ops.so> [iterations]\n", argv[0]);
    return 2;
  }
  api ours{}, ref{};
  if (!load(argv[1], &ours) || !load(argv[2], &ref)) return 2;
  decltype(&fork_get_device_count) devcount;
  decltype(&wholememory_init) init_fn;
  if (!sym(ours.so, "fork_get_device_count", &devcount) || !sym(ours.so, "wholememory_init", &init_fn)) return 2;
  if (devcount() != 0) {
    printf("ops_validation_diff: a GPU is present, nothing compared (the equivalence holds on a GPU-less machine only)\n");
    return 0;
  }
  init_fn(0, LEVEL_FATAL);
  const long iters = argc > 3 ? atol(argv[3]) : 100000;
  std::mt19937_64 rng(77);
  auto pick = [&](long lo, long hi) { return lo + (long)(rng() % (uint64_t)(hi - lo + 1)); };
  static char table_mem[1 << 16], idx_mem[1 << 12], dense_mem[1 << 16];
  const wholememory_dtype_t dts[] = {WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_HALF, WHOLEMEMORY_DT_INT, WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_INT8};
  int divergences = 0;
  long dispatched = 0, refused = 0, superset = 0;

  for (long it = 0; it < iters; ++it) {
    const int64_t n = pick(0, 20);
    auto td = make_desc((int)pick(1, 2), pick(1, 30), pick(1, 16), dts[pick(0, 4)]);
    auto id = make_desc((int)pick(1, 2), n, pick(1, 3), pick(0, 3) == 0 ? WHOLEMEMORY_DT_FLOAT : (pick(0, 1) ? WHOLEMEMORY_DT_INT : WHOLEMEMORY_DT_INT64));
    auto dd = make_desc((int)pick(1, 2), pick(0, 2) == 0 ? pick(0, 20) : n, pick(0, 3) == 0 ? pick(1, 16) : td.sizes[1], dts[pick(0, 4)]);
    /* one corruption, applied to the live description of one operand after creation */
    const int corrupt_which = (int)pick(0, 5); /* 0..2: table / indices / dense, else none */
    const int corrupt_how   = (int)pick(0, 3);
    int codes[2][2];
    for (int side = 0; side < 2; ++side) {
      api& a = side == 0 ? ours : ref;
      wholememory_tensor_t t = nullptr, i = nullptr, d = nullptr;
      wholememory_tensor_description_t td_ = td, id_ = id, dd_ = dd;
      if (a.from_ptr(&t, table_mem, &td_) != WHOLEMEMORY_SUCCESS || a.from_ptr(&i, idx_mem, &id_) != WHOLEMEMORY_SUCCESS ||
          a.from_ptr(&d, dense_mem, &dd_) != WHOLEMEMORY_SUCCESS) {
        fprintf(stderr, "setup failed on %s\n", side == 0 ? "ours" : "reference");
        return 2;
      }
      if (corrupt_which < 3) {
        wholememory_tensor_description_t* live = a.get_desc(corrupt_which == 0 ? t : corrupt_which == 1 ? i : d);
        switch (corrupt_how) {
          case 0: live->dim = 3; break;
          case 1: live->strides[live->dim - 1] = 2; break;
          case 2: live->dtype = WHOLEMEMORY_DT_UNKNOWN; break;
          default: live->dim = 0; break;
        }
      }
      codes[side][0] = (int)a.gather(t, i, d, nullptr, nullptr, -1);
      codes[side][1] = (int)a.scatter(d, i, t, nullptr, nullptr, -1);
      a.destroy(t);
      a.destroy(i);
      a.destroy(d);
    }
    for (int op = 0; op < 2; ++op) {
      const int mine = codes[0][op], theirs = codes[1][op];
      bool same = theirs == 1000 ? mine == (int)WHOLEMEMORY_CUDA_ERROR : mine == theirs;
      /* the one documented superset: a 1-D table with a 1-D dense operand.  The reference compares
       * ranks after unsqueezing the table and refuses it (INVALID_INPUT); this library accepts it next to the [n, 1] form. */
      const int eff_table_dim = (corrupt_which == 0 && corrupt_how == 0) ? 3 : (corrupt_which == 0 && corrupt_how == 3) ? 0 : td.dim;
      const int eff_dense_dim = (corrupt_which == 2 && corrupt_how == 0) ? 3 : (corrupt_which == 2 && corrupt_how == 3) ? 0 : dd.dim;
      if (!same && eff_table_dim == 1 && eff_dense_dim == 1 && theirs == (int)WHOLEMEMORY_INVALID_INPUT && mine == (int)WHOLEMEMORY_CUDA_ERROR) {
        same = true;
        ++superset;
      }
      (theirs == 1000 ? dispatched : refused)++;
      if (!same && ++divergences <= 40)
        fprintf(stderr, "DIVERGENCE [%s] iteration %ld: ours %d reference %d | table %s | indices %s | dense %s | corrupt %d/%d\n",
                op == 0 ? "gather" : "scatter", it, mine, theirs, show(td).c_str(), show(id).c_str(), show(dd).c_str(), corrupt_which, corrupt_how);
    }
  }
  printf("ops_validation_diff: %ld iterations, %d divergences (%ld calls dispatched by the reference, %ld refused, %ld of those the documented 1-D/1-D superset)\n",
         iters, divergences, dispatched, refused, superset);
  return divergences > 100 ? 100 : divergences;
}
