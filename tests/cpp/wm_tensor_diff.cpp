/*
 * Differential test of the tensor <-> partition arithmetic for HANDLE-BACKED tensors: entry offsets, entry partition
 * sizes, local entry count / start, local-tensor mapping, data pointers -- the reference's wholememory_tensor.cpp (compiled
 * for the CPU into oracle/_ref/ref_host_tensor.so) against this repo's libwholegraph.so.
 *
 * A real WholeMemory handle cannot exist without a GPU, so the handle is SYNTHETIC: this program is compiled with this
 * repo's internal header and fills a wholememory_handle_ / wholememory_comm_ pair by hand (world size, rank, memory type,
 * byte partition, fake base addresses).  This repo's tensor code reads those fields directly; the reference's tensor code
 * reads them through the C-ABI accessors (wholememory_get_local_memory, ..._rank_partition_offsets, ...), which resolve
 * to this repo's library and therefore see the same synthetic handle.  Both sides then answer the same questions.
 *
 *   wm_tensor_diff <ref_host_tensor.so> [iterations]      (links libwholegraph.so directly)   exit code = divergences (capped)
 */
#include "wm_internal.hpp"

#include <dlfcn.h>

#include <random>
#include <string>

namespace {

struct ref_api {
  decltype(&wholememory_make_tensor_from_handle) from_handle;
  decltype(&wholememory_destroy_tensor) destroy;
  decltype(&wholememory_tensor_get_tensor_description) get_desc;
  decltype(&wholememory_tensor_get_subtensor) subtensor;
  decltype(&wholememory_tensor_get_entry_offsets) entry_offsets;
  decltype(&wholememory_tensor_get_entry_partition_sizes) partition_sizes;
  decltype(&wholememory_tensor_get_local_entry_count) local_count;
  decltype(&wholememory_tensor_get_local_entry_start) local_start;
  decltype(&wholememory_tensor_map_local_tensor) map_local;
  decltype(&wholememory_tensor_get_data_pointer) data_ptr;
  decltype(&wholememory_tensor_has_handle) has_handle;
  decltype(&wholememory_tensor_get_memory_handle) get_handle;
};
template <typename F>
bool sym(void* so, const char* name, F* out)
{
  *out = reinterpret_cast<F>(dlsym(so, name));
  if (*out == nullptr) fprintf(stderr, "missing %s: %s\n", name, dlerror());
  return *out != nullptr;
}

int g_div = 0;
void diverge(const char* what, long it, const std::string& detail)
{
  if (++g_div <= 40) fprintf(stderr, "DIVERGENCE [%s] iteration %ld: %s\n", what, it, detail.c_str());
}
std::string show(const wholememory_tensor_description_t& d)
{
  return "dim " + std::to_string(d.dim) + " dtype " + std::to_string((int)d.dtype) + " off " + std::to_string(d.storage_offset) + " sizes " +
         std::to_string(d.sizes[0]) + "," + std::to_string(d.sizes[1]) + " strides " + std::to_string(d.strides[0]) + "," + std::to_string(d.strides[1]);
}
bool same_desc(const wholememory_tensor_description_t& x, const wholememory_tensor_description_t& y)
{
  if (x.dim != y.dim || x.dtype != y.dtype || x.storage_offset != y.storage_offset) return false;
  for (int i = 0; i < x.dim; ++i)
    if (x.sizes[i] != y.sizes[i] || x.strides[i] != y.strides[i]) return false;
  return true;
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc < 2) {
    fprintf(stderr, "usage: %s <ref_host_tensor.so> [iterations]\n", argv[0]);
    return 2;
  }
  void* so = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!so) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 2;
  }
  ref_api ref{};
  if (!(sym(so, "wholememory_make_tensor_from_handle", &ref.from_handle) && sym(so, "wholememory_destroy_tensor", &ref.destroy) &&
        sym(so, "wholememory_tensor_get_tensor_description", &ref.get_desc) && sym(so, "wholememory_tensor_get_subtensor", &ref.subtensor) &&
        sym(so, "wholememory_tensor_get_entry_offsets", &ref.entry_offsets) &&
        sym(so, "wholememory_tensor_get_entry_partition_sizes", &ref.partition_sizes) &&
        sym(so, "wholememory_tensor_get_local_entry_count", &ref.local_count) &&
        sym(so, "wholememory_tensor_get_local_entry_start", &ref.local_start) && sym(so, "wholememory_tensor_map_local_tensor", &ref.map_local) &&
        sym(so, "wholememory_tensor_get_data_pointer", &ref.data_ptr) && sym(so, "wholememory_tensor_has_handle", &ref.has_handle) &&
        sym(so, "wholememory_tensor_get_memory_handle", &ref.get_handle)))
    return 2;
  const long iters = argc > 2 ? atol(argv[2]) : 100000;
  std::mt19937_64 rng(909);
  auto pick = [&](long lo, long hi) { return lo + (long)(rng() % (uint64_t)(hi - lo + 1)); };
  static char arena[1 << 20]; /* addresses only; never dereferenced */
  const wholememory_dtype_t dts[] = {WHOLEMEMORY_DT_FLOAT, WHOLEMEMORY_DT_HALF, WHOLEMEMORY_DT_INT64, WHOLEMEMORY_DT_INT8, WHOLEMEMORY_DT_DOUBLE};
  long truncated_beyond = 0;

  for (long it = 0; it < iters; ++it) {
    /* ---- a synthetic row-partitioned allocation */
    const int ws            = (int)pick(1, 8);
    const int me            = (int)pick(0, ws - 1);
    const int dim           = (int)pick(1, 2);
    const auto dt           = dts[pick(0, 4)];
    const size_t esz        = wholememory_dtype_get_element_size(dt);
    const int64_t cols      = dim == 2 ? pick(1, 20) : 1;
    const int64_t stride    = dim == 2 ? cols + (pick(0, 2) == 0 ? pick(1, 4) : 0) : 1;
    const int64_t rows      = pick(ws, 400);
    const size_t row_bytes  = (size_t)stride * esz;
    wholememory_comm_ comm;
    comm.world_rank = me, comm.world_size = ws;
    wholememory_handle_ h;
    h.comm        = &comm;
    h.type        = (wholememory_memory_type_t)pick(1, 3);
    h.location    = WHOLEMEMORY_ML_HOST;
    h.granularity = row_bytes;
    h.total_size  = (size_t)rows * row_bytes;
    h.part_sizes.assign(ws, 0);
    h.part_offsets.assign(ws + 1, 0);
    if (pick(0, 1) == 0) { /* equal split: ceil(rows / ws) rows per rank, tail ranks short or empty */
      size_t per = ((size_t)rows + ws - 1) / ws;
      for (int r = 0; r < ws; ++r) {
        size_t b = std::min((size_t)r * per, (size_t)rows), e = std::min((size_t)(r + 1) * per, (size_t)rows);
        h.part_offsets[r] = b * row_bytes, h.part_sizes[r] = (e - b) * row_bytes;
      }
    } else { /* custom split: at least one row each */
      std::vector<size_t> cnt(ws, 1);
      for (int64_t left = rows - ws; left > 0; --left) cnt[pick(0, ws - 1)]++;
      size_t acc = 0;
      for (int r = 0; r < ws; ++r) h.part_offsets[r] = acc * row_bytes, h.part_sizes[r] = cnt[r] * row_bytes, acc += cnt[r];
    }
    h.part_offsets[ws] = (size_t)rows * row_bytes;
    h.rank_base.assign(ws, nullptr);
    for (int r = 0; r < ws; ++r) h.rank_base[r] = arena + h.part_offsets[r];
    h.local_ptr   = h.part_sizes[me] ? h.rank_base[me] : nullptr;
    h.flat_base   = h.type == WHOLEMEMORY_MT_CONTINUOUS ? arena : nullptr;
    h.peer_mapped = true;
    /* hand-made objects: tell the library's live-object registry about them (every entry point checks it), and take
     * them out again when this iteration's scope ends */
    struct registered {
      const void *c, *h;
      registered(const void* c_, const void* h_) : c(c_), h(h_)
      {
        wm::obj_register(wm::OBJ_COMM, c);
        wm::obj_register(wm::OBJ_HANDLE, h);
      }
      ~registered()
      {
        wm::obj_unregister(wm::OBJ_HANDLE, h);
        wm::obj_unregister(wm::OBJ_COMM, c);
      }
    } reg(&comm, &h);

    wholememory_tensor_description_t d;
    wholememory_initialize_tensor_desc(&d);
    d.dim = dim, d.dtype = dt, d.storage_offset = 0;
    d.sizes[0] = rows, d.strides[0] = dim == 2 ? stride : 1;
    if (dim == 2) d.sizes[1] = cols, d.strides[1] = 1;

    wholememory_tensor_t t1 = nullptr, t2 = nullptr;
    wholememory_tensor_description_t d1 = d, d2 = d;
    auto e1 = wholememory_make_tensor_from_handle(&t1, &h, &d1);
    auto e2 = ref.from_handle(&t2, &h, &d2);
    if ((int)e1 != (int)e2) diverge("make_tensor_from_handle", it, show(d));
    if (e1 != WHOLEMEMORY_SUCCESS || e2 != WHOLEMEMORY_SUCCESS) continue;

    /* optionally a view that starts at row 0 / column c0 and may drop trailing rows and columns */
    wholememory_tensor_t v1 = t1, v2 = t2;
    bool is_view = pick(0, 1) == 0;
    int64_t keep_rows = rows;
    if (is_view) {
      int64_t st[2] = {0, dim == 2 ? pick(0, cols - 1) : 0};
      keep_rows     = pick(0, 3) == 0 ? pick(1, rows) : rows;
      int64_t en[2] = {keep_rows, dim == 2 ? pick(st[1] + 1, cols) : 0};
      auto r1 = wholememory_tensor_get_subtensor(t1, st, en, &v1), r2 = ref.subtensor(t2, st, en, &v2);
      if ((int)r1 != (int)r2 || r1 != WHOLEMEMORY_SUCCESS) {
        if ((int)r1 != (int)r2) diverge("get_subtensor", it, show(d));
        wholememory_destroy_tensor(t1);
        ref.destroy(t2);
        continue;
      }
      if (!same_desc(*wholememory_tensor_get_tensor_description(v1), *ref.get_desc(v2))) diverge("get_subtensor (description)", it, show(d));
    }

    std::vector<size_t> a(ws + 1, 0), b(ws + 1, 0);
    auto r1 = wholememory_tensor_get_entry_offsets(a.data(), v1), r2 = ref.entry_offsets(b.data(), v2);
    if ((int)r1 != (int)r2 || a != b) diverge("get_entry_offsets", it, show(d) + " ws " + std::to_string(ws));
    a.assign(ws + 1, 0), b.assign(ws + 1, 0);
    r1 = wholememory_tensor_get_entry_partition_sizes(a.data(), v1), r2 = ref.partition_sizes(b.data(), v2);
    if ((int)r1 != (int)r2 || a != b) diverge("get_entry_partition_sizes", it, show(d) + " ws " + std::to_string(ws));
    size_t c1 = 0, c2 = 0;
    r1 = wholememory_tensor_get_local_entry_count(&c1, v1), r2 = ref.local_count(&c2, v2);
    if ((int)r1 != (int)r2 || c1 != c2) diverge("get_local_entry_count", it, std::to_string(c1) + " vs " + std::to_string(c2));
    r1 = wholememory_tensor_get_local_entry_start(&c1, v1), r2 = ref.local_start(&c2, v2);
    if ((int)r1 != (int)r2 || c1 != c2) diverge("get_local_entry_start", it, std::to_string(c1) + " vs " + std::to_string(c2));
    if ((char*)wholememory_tensor_get_data_pointer(v1) != (char*)ref.data_ptr(v2)) diverge("get_data_pointer", it, show(d));
    if (wholememory_tensor_has_handle(v1) != ref.has_handle(v2) || wholememory_tensor_get_memory_handle(v1) != ref.get_handle(v2))
      diverge("has_handle / get_memory_handle", it, show(d));

    /* local mapping.  One documented difference: when a row-truncated view ends BEFORE this rank's partition begins, the
     * reference's unsigned subtraction wraps (wholememory_tensor.cpp:262) and it reports the whole local partition; this
     * library reports an empty local tensor.  Those cases are counted, not compared. */
    const bool beyond = (size_t)keep_rows * row_bytes < h.part_offsets[me];
    wholememory_tensor_t l1 = nullptr, l2 = nullptr;
    r1 = wholememory_tensor_map_local_tensor(v1, &l1), r2 = ref.map_local(v2, &l2);
    if (beyond) {
      ++truncated_beyond;
    } else {
      if ((int)r1 != (int)r2) diverge("map_local_tensor (code)", it, show(*wholememory_tensor_get_tensor_description(v1)) + " ours " + std::to_string((int)r1) + " reference " + std::to_string((int)r2));
      if (r1 == WHOLEMEMORY_SUCCESS && r2 == WHOLEMEMORY_SUCCESS) {
        if (!same_desc(*wholememory_tensor_get_tensor_description(l1), *ref.get_desc(l2)))
          diverge("map_local_tensor (description)", it,
                  show(*wholememory_tensor_get_tensor_description(l1)) + " vs " + show(*ref.get_desc(l2)) + " rank " + std::to_string(me) + "/" + std::to_string(ws));
        if ((char*)wholememory_tensor_get_data_pointer(l1) != (char*)ref.data_ptr(l2)) diverge("map_local_tensor (pointer)", it, show(d));
      }
    }
    if (l1) wholememory_destroy_tensor(l1);
    if (l2) ref.destroy(l2);
    if (is_view) {
      wholememory_destroy_tensor(v1);
      ref.destroy(v2);
    }
    wholememory_destroy_tensor(t1);
    ref.destroy(t2);
  }
  printf("wm_tensor_diff: %ld iterations, %d divergences (%ld truncated-view cases beyond the local partition not compared)\n", iters, g_div,
         truncated_beyond);
  return g_div > 100 ? 100 : g_div;
}
