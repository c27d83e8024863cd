/*
 * Differential test of the HOST-side descriptor and view logic: the reference's own tensor_description.cpp and
 * wholememory_tensor.cpp (compiled for the CPU from /root/reference into oracle/_ref/ref_host_tensor.so by
 * oracle/build_ref_host_tensor.sh) against this repo's libwholegraph.so, function by function on the same randomised
 * inputs.  Both libraries are dlopen'ed RTLD_LOCAL and every entry point is looked up per handle, so equal symbol names do
 * not interfere.  Inputs stay inside what the reference tolerates (it has no bounds checks on a few paths and would read
 * out of range); everything else -- including invalid descriptors that must be REFUSED -- is compared exactly:
 * return values, error codes, resulting descriptors byte for byte, data pointers.
 *
 *   host_diff_test <ours.so> <reference_host.so> [iterations]        exit code = number of divergences (capped)
 */
#include <wholememory/tensor_description.h>
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
  decltype(&wholememory_dtype_get_element_size) dtype_size;
  decltype(&wholememory_dtype_is_floating_number) is_float;
  decltype(&wholememory_dtype_is_integer_number) is_int;
  decltype(&wholememory_create_array_desc) create_array;
  decltype(&wholememory_create_matrix_desc) create_matrix;
  decltype(&wholememory_initialize_tensor_desc) init_tensor;
  decltype(&wholememory_copy_array_desc_to_matrix) a2m;
  decltype(&wholememory_copy_array_desc_to_tensor) a2t;
  decltype(&wholememory_copy_matrix_desc_to_tensor) m2t;
  decltype(&wholememory_convert_tensor_desc_to_array) t2a;
  decltype(&wholememory_convert_tensor_desc_to_matrix) t2m;
  decltype(&wholememory_get_memory_element_count_from_array) cnt_a;
  decltype(&wholememory_get_memory_size_from_array) size_a;
  decltype(&wholememory_get_memory_element_count_from_matrix) cnt_m;
  decltype(&wholememory_get_memory_size_from_matrix) size_m;
  decltype(&wholememory_get_memory_element_count_from_tensor) cnt_t;
  decltype(&wholememory_get_memory_size_from_tensor) size_t_;
  decltype(&wholememory_squeeze_tensor) squeeze;
  decltype(&wholememory_unsqueeze_tensor) unsqueeze;
  decltype(&wholememory_make_tensor_from_pointer) from_ptr;
  decltype(&wholememory_destroy_tensor) destroy;
  decltype(&wholememory_tensor_has_handle) has_handle;
  decltype(&wholememory_tensor_get_tensor_description) get_desc;
  decltype(&wholememory_tensor_get_data_pointer) data_ptr;
  decltype(&wholememory_tensor_get_subtensor) subtensor;
  decltype(&wholememory_tensor_get_root) root;
  decltype(&wholememory_create_tensor) create_tensor;
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
  return sym(a->so, "wholememory_dtype_get_element_size", &a->dtype_size) && sym(a->so, "wholememory_dtype_is_floating_number", &a->is_float) &&
         sym(a->so, "wholememory_dtype_is_integer_number", &a->is_int) && sym(a->so, "wholememory_create_array_desc", &a->create_array) &&
         sym(a->so, "wholememory_create_matrix_desc", &a->create_matrix) && sym(a->so, "wholememory_initialize_tensor_desc", &a->init_tensor) &&
         sym(a->so, "wholememory_copy_array_desc_to_matrix", &a->a2m) && sym(a->so, "wholememory_copy_array_desc_to_tensor", &a->a2t) &&
         sym(a->so, "wholememory_copy_matrix_desc_to_tensor", &a->m2t) && sym(a->so, "wholememory_convert_tensor_desc_to_array", &a->t2a) &&
         sym(a->so, "wholememory_convert_tensor_desc_to_matrix", &a->t2m) &&
         sym(a->so, "wholememory_get_memory_element_count_from_array", &a->cnt_a) && sym(a->so, "wholememory_get_memory_size_from_array", &a->size_a) &&
         sym(a->so, "wholememory_get_memory_element_count_from_matrix", &a->cnt_m) &&
         sym(a->so, "wholememory_get_memory_size_from_matrix", &a->size_m) &&
         sym(a->so, "wholememory_get_memory_element_count_from_tensor", &a->cnt_t) &&
         sym(a->so, "wholememory_get_memory_size_from_tensor", &a->size_t_) && sym(a->so, "wholememory_squeeze_tensor", &a->squeeze) &&
         sym(a->so, "wholememory_unsqueeze_tensor", &a->unsqueeze) && sym(a->so, "wholememory_make_tensor_from_pointer", &a->from_ptr) &&
         sym(a->so, "wholememory_destroy_tensor", &a->destroy) && sym(a->so, "wholememory_tensor_has_handle", &a->has_handle) &&
         sym(a->so, "wholememory_tensor_get_tensor_description", &a->get_desc) && sym(a->so, "wholememory_tensor_get_data_pointer", &a->data_ptr) &&
         sym(a->so, "wholememory_tensor_get_subtensor", &a->subtensor) && sym(a->so, "wholememory_tensor_get_root", &a->root) &&
         sym(a->so, "wholememory_create_tensor", &a->create_tensor);
}

int g_div = 0;
void diverge(const char* what, long iter, const std::string& detail)
{
  if (++g_div <= 40) fprintf(stderr, "DIVERGENCE [%s] iteration %ld: %s\n", what, iter, detail.c_str());
}
#define SAME(what, a, b)                                                                                   \
  do {                                                                                                     \
    auto va_ = (a);                                                                                        \
    auto vb_ = (b);                                                                                        \
    if (va_ != vb_) diverge(what, it, std::string("ours ") + std::to_string((long long)va_) + " reference " + std::to_string((long long)vb_)); \
  } while (0)

/* compare only the fields the ABI defines as meaningful: sizes/strides of the first `dim` dims, offset, dim, dtype */
bool same_desc(const wholememory_tensor_description_t& x, const wholememory_tensor_description_t& y)
{
  if (x.dim != y.dim || x.dtype != y.dtype || x.storage_offset != y.storage_offset) return false;
  for (int i = 0; i < x.dim && i < WHOLEMEMORY_MAX_TENSOR_DIM; ++i)
    if (x.sizes[i] != y.sizes[i] || x.strides[i] != y.strides[i]) return false;
  return true;
}
std::string show(const wholememory_tensor_description_t& d)
{
  std::string s = "dim " + std::to_string(d.dim) + " dtype " + std::to_string((int)d.dtype) + " off " + std::to_string(d.storage_offset) + " sizes";
  for (int i = 0; i < d.dim && i < 8; ++i) s += " " + std::to_string(d.sizes[i]);
  s += " strides";
  for (int i = 0; i < d.dim && i < 8; ++i) s += " " + std::to_string(d.strides[i]);
  return s;
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s <ours.so> <reference_host.so> [iterations]\n", argv[0]);
    return 2;
  }
  api ours{}, ref{};
  if (!load(argv[1], &ours) || !load(argv[2], &ref)) return 2;
  const long iters = argc > 3 ? atol(argv[3]) : 200000;
  std::mt19937_64 rng(20241017);
  auto pick = [&](long lo, long hi) { return lo + (long)(rng() % (uint64_t)(hi - lo + 1)); };
  static char arena[1 << 16];

  for (int dt = -2; dt <= 12; ++dt) {
    long it = dt;
    SAME("dtype_get_element_size", (long long)ours.dtype_size((wholememory_dtype_t)dt), (long long)ref.dtype_size((wholememory_dtype_t)dt));
    SAME("dtype_is_floating_number", ours.is_float((wholememory_dtype_t)dt), ref.is_float((wholememory_dtype_t)dt));
    SAME("dtype_is_integer_number", ours.is_int((wholememory_dtype_t)dt), ref.is_int((wholememory_dtype_t)dt));
  }

  for (long it = 0; it < iters; ++it) {
    /* ---- a random tensor description: mostly plausible, sometimes broken in one field */
    wholememory_tensor_description_t d;
    ours.init_tensor(&d);
    {
      wholememory_tensor_description_t dr;
      memset(&dr, 0x5a, sizeof(dr));
      ref.init_tensor(&dr);
      wholememory_tensor_description_t dz = d;
      if (memcmp(&dz, &dr, sizeof(dz)) != 0 && it == 0) diverge("initialize_tensor_desc", it, show(dz) + " vs " + show(dr));
    }
    d.dim            = (int)pick(0, 4);
    d.dtype          = (wholememory_dtype_t)pick(0, 9);
    d.storage_offset = pick(0, 3) == 0 ? pick(0, 500) : 0;
    int64_t stride   = 1;
    for (int i = d.dim - 1; i >= 0; --i) {
      d.sizes[i]   = pick(0, 5) == 0 ? 1 : pick(1, 40);
      d.strides[i] = stride + (pick(0, 6) == 0 ? pick(1, 5) : 0);
      stride       = d.strides[i] * d.sizes[i];
    }
    if (d.dim > 0 && pick(0, 9) == 0) d.strides[d.dim - 1] = pick(0, 3);

    /* conversions */
    {
      wholememory_array_description_t a1, a2;
      memset(&a1, 0, sizeof(a1));
      memset(&a2, 0, sizeof(a2));
      wholememory_tensor_description_t t1 = d, t2 = d;
      bool r1 = ours.t2a(&a1, &t1), r2 = ref.t2a(&a2, &t2);
      SAME("convert_tensor_desc_to_array", r1, r2);
      if (r1 && r2 && (a1.size != a2.size || a1.storage_offset != a2.storage_offset || a1.dtype != a2.dtype))
        diverge("convert_tensor_desc_to_array (fields)", it, show(d));
      wholememory_matrix_description_t m1, m2;
      memset(&m1, 0, sizeof(m1));
      memset(&m2, 0, sizeof(m2));
      r1 = ours.t2m(&m1, &t1), r2 = ref.t2m(&m2, &t2);
      SAME("convert_tensor_desc_to_matrix", r1, r2);
      if (r1 && r2 && (m1.sizes[0] != m2.sizes[0] || m1.sizes[1] != m2.sizes[1] || m1.stride != m2.stride ||
                       m1.storage_offset != m2.storage_offset || m1.dtype != m2.dtype))
        diverge("convert_tensor_desc_to_matrix (fields)", it, show(d));
      if (r1 && r2) {
        SAME("get_memory_element_count_from_matrix", ours.cnt_m(&m1), ref.cnt_m(&m2));
        SAME("get_memory_size_from_matrix", ours.size_m(&m1), ref.size_m(&m2));
        wholememory_tensor_description_t b1, b2;
        ours.m2t(&b1, &m1);
        ref.m2t(&b2, &m2);
        if (!same_desc(b1, b2)) diverge("copy_matrix_desc_to_tensor", it, show(b1) + " vs " + show(b2));
      }
      wholememory_array_description_t ca1 = ours.create_array(d.sizes[0], d.storage_offset, d.dtype);
      wholememory_array_description_t ca2 = ref.create_array(d.sizes[0], d.storage_offset, d.dtype);
      if (ca1.size != ca2.size || ca1.storage_offset != ca2.storage_offset || ca1.dtype != ca2.dtype) diverge("create_array_desc", it, show(d));
      SAME("get_memory_element_count_from_array", ours.cnt_a(&ca1), ref.cnt_a(&ca2));
      SAME("get_memory_size_from_array", ours.size_a(&ca1), ref.size_a(&ca2));
      wholememory_matrix_description_t am1, am2;
      ours.a2m(&am1, &ca1);
      ref.a2m(&am2, &ca2);
      if (am1.sizes[0] != am2.sizes[0] || am1.sizes[1] != am2.sizes[1] || am1.stride != am2.stride || am1.storage_offset != am2.storage_offset ||
          am1.dtype != am2.dtype)
        diverge("copy_array_desc_to_matrix", it, show(d));
      wholememory_tensor_description_t at1, at2;
      ours.a2t(&at1, &ca1);
      ref.a2t(&at2, &ca2);
      if (!same_desc(at1, at2)) diverge("copy_array_desc_to_tensor", it, show(at1) + " vs " + show(at2));
      int64_t sz[2] = {d.sizes[0], d.sizes[1]};
      auto cm1      = ours.create_matrix(sz, d.strides[0], d.storage_offset, d.dtype);
      auto cm2      = ref.create_matrix(sz, d.strides[0], d.storage_offset, d.dtype);
      if (cm1.sizes[0] != cm2.sizes[0] || cm1.sizes[1] != cm2.sizes[1] || cm1.stride != cm2.stride || cm1.storage_offset != cm2.storage_offset ||
          cm1.dtype != cm2.dtype)
        diverge("create_matrix_desc", it, show(d));
    }
    SAME("get_memory_element_count_from_tensor", ours.cnt_t(&d), ref.cnt_t(&d));
    SAME("get_memory_size_from_tensor", ours.size_t_(&d), ref.size_t_(&d));

    /* squeeze / unsqueeze (the reference indexes strides[dim - 1] and sizes[dim]: keep 1 <= dim <= 6) */
    if (d.dim >= 1 && d.dim <= 6) {
      int at                              = (int)pick(-1, d.dim + 1);
      wholememory_tensor_description_t s1 = d, s2 = d;
      bool r1 = ours.squeeze(&s1, at), r2 = ref.squeeze(&s2, at);
      SAME("squeeze_tensor", r1, r2);
      if (r1 && r2 && !same_desc(s1, s2)) diverge("squeeze_tensor (result)", it, show(d) + " at " + std::to_string(at));
      s1 = d, s2 = d;
      r1 = ours.unsqueeze(&s1, at), r2 = ref.unsqueeze(&s2, at);
      SAME("unsqueeze_tensor", r1, r2);
      if (r1 && r2 && !same_desc(s1, s2))
        diverge("unsqueeze_tensor (result)", it, show(d) + " at " + std::to_string(at) + " -> " + show(s1) + " vs " + show(s2));
    }

    /* pointer tensors and sub-tensors */
    {
      wholememory_tensor_t t1 = nullptr, t2 = nullptr;
      wholememory_tensor_description_t d1 = d, d2 = d;
      auto e1 = ours.from_ptr(&t1, arena, &d1), e2 = ref.from_ptr(&t2, arena, &d2);
      SAME("make_tensor_from_pointer", (int)e1, (int)e2);
      if (e1 == WHOLEMEMORY_SUCCESS && e2 == WHOLEMEMORY_SUCCESS) {
        SAME("tensor_has_handle", ours.has_handle(t1), ref.has_handle(t2));
        if (!same_desc(*ours.get_desc(t1), *ref.get_desc(t2))) diverge("tensor_get_tensor_description", it, show(d));
        SAME("tensor_get_data_pointer", (long long)((char*)ours.data_ptr(t1) - arena), (long long)((char*)ref.data_ptr(t2) - arena));
        for (int rep = 0; rep < 3; ++rep) {
          int64_t st[2], en[2];
          for (int i = 0; i < 2; ++i) {
            int64_t size = i < d.dim ? d.sizes[i] : 1;
            st[i]        = pick(0, 5) == 0 ? -1 : pick(-1, size + 1);
            /* the reference does not bound `ends` from above (wholememory_tensor.cpp:430-442): keep ends <= size so that
             * both sides describe memory that exists; everything else (empty, reversed, negative, -1) is compared */
            en[i] = pick(0, 5) == 0 ? -1 : pick(-2, size);
          }
          wholememory_tensor_t s1 = nullptr, s2 = nullptr;
          auto r1 = ours.subtensor(t1, st, en, &s1), r2 = ref.subtensor(t2, st, en, &s2);
          if ((int)r1 != (int)r2)
            diverge("tensor_get_subtensor (code)", it,
                    show(d) + " starts " + std::to_string(st[0]) + "," + std::to_string(st[1]) + " ends " + std::to_string(en[0]) + "," +
                      std::to_string(en[1]) + " ours " + std::to_string((int)r1) + " reference " + std::to_string((int)r2));
          if (r1 == WHOLEMEMORY_SUCCESS && r2 == WHOLEMEMORY_SUCCESS) {
            if (!same_desc(*ours.get_desc(s1), *ref.get_desc(s2)))
              diverge("tensor_get_subtensor (description)", it, show(*ours.get_desc(s1)) + " vs " + show(*ref.get_desc(s2)));
            SAME("sub-tensor data pointer", (long long)((char*)ours.data_ptr(s1) - arena), (long long)((char*)ref.data_ptr(s2) - arena));
            if (ours.root(s1) != t1 || ref.root(s2) != t2) diverge("tensor_get_root", it, show(d));
          }
          if (s1) ours.destroy(s1);
          if (s2) ref.destroy(s2);
        }
      }
      if (t1) ours.destroy(t1);
      if (t2) ref.destroy(t2);
    }
  }
  long create_cases = 0;
  /* ---- wholememory_create_tensor: argument checks and the error code handed back when the allocation itself cannot
   * succeed.  The communicator comes from this repo's library (the reference code only touches it through the C ABI);
   * without a GPU every well-formed request ends in the same wholememory_malloc refusal on both sides, with a GPU the
   * section is skipped (it would allocate). */
  {
    decltype(&wholememory_init) init_fn;
    decltype(&wholememory_finalize) fini_fn;
    decltype(&wholememory_create_unique_id) uid_fn;
    decltype(&wholememory_create_communicator) comm_fn;
    decltype(&wholememory_destroy_communicator) comm_free_fn;
    decltype(&fork_get_device_count) devcount_fn;
    if (sym(ours.so, "wholememory_init", &init_fn) && sym(ours.so, "wholememory_finalize", &fini_fn) &&
        sym(ours.so, "wholememory_create_unique_id", &uid_fn) && sym(ours.so, "wholememory_create_communicator", &comm_fn) &&
        sym(ours.so, "wholememory_destroy_communicator", &comm_free_fn) && sym(ours.so, "fork_get_device_count", &devcount_fn) &&
        devcount_fn() == 0) {
      init_fn(0, LEVEL_FATAL);
      wholememory_unique_id_t uid;
      wholememory_comm_t comm = nullptr;
      if (uid_fn(&uid) == WHOLEMEMORY_SUCCESS && comm_fn(&comm, uid, 0, 1) == WHOLEMEMORY_SUCCESS) {
        for (long it = 0; it < 4000; ++it) {
          wholememory_tensor_description_t d;
          ours.init_tensor(&d);
          d.dim            = (int)pick(0, 3);
          d.dtype          = (wholememory_dtype_t)pick(0, 9);
          d.storage_offset = pick(0, 4) == 0 ? pick(1, 9) : 0;
          d.sizes[0] = pick(1, 50), d.sizes[1] = pick(1, 20);
          d.strides[1] = pick(0, 6) == 0 ? 2 : 1;
          d.strides[0] = d.dim == 2 ? d.sizes[1] * d.strides[1] + pick(0, 3) : (pick(0, 6) == 0 ? 2 : 1);
          auto mt  = (wholememory_memory_type_t)pick(1, 3);
          auto loc = (wholememory_memory_location_t)pick(1, 2);
          wholememory_tensor_t t1 = nullptr, t2 = nullptr;
          wholememory_tensor_description_t d1 = d, d2 = d;
          auto r1 = ours.create_tensor(&t1, &d1, comm, mt, loc, nullptr);
          auto r2 = ref.create_tensor(&t2, &d2, comm, mt, loc, nullptr);
          ++create_cases;
          if ((int)r1 != (int)r2) diverge("create_tensor (code)", it, show(d) + " ours " + std::to_string((int)r1) + " reference " + std::to_string((int)r2));
          if (r1 == WHOLEMEMORY_SUCCESS && t1) ours.destroy(t1);
          if (r2 == WHOLEMEMORY_SUCCESS && t2) ref.destroy(t2);
        }
        comm_free_fn(comm);
      }
      fini_fn();
    }
  }
  printf("host_diff_test: %ld iterations, %d divergences (%ld create_tensor cases)\n", iters, g_div, create_cases);
  return g_div > 100 ? 100 : g_div;
}
