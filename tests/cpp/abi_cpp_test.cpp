/*
 * C++ caller's view of the drop-in boundary: a plain C++17 program that includes ONLY the headers under
 * include/wholememory and links libwholegraph.so, the way the reference's own C++ tests and bench do (the .cu suites
 * under cpp/tests/wholememory_ops and cpp/bench/wholememory_ops/gather_scatter_bench.cu link wholegraph::wholegraph).
 * It exercises the C entry points with their C++ default arguments, the C++-only helpers (wholememory::get_default_env_func / get_cached_env_func) and
 * caller-supplied env functions with a caller-defined memory context, like the reference tests' output contexts.
 *
 *   abi_cpp_test cpu          host-only checks (no GPU needed): descriptors, views, communicators (incl. 2 forked ranks),
 *                             env-function protocol, and "device ops fail loudly without a GPU"
 *   abi_cpp_test gpu [ranks]  gather / scatter / SGD step / neighbor sampling on cuda:0, `ranks` forked processes sharing
 *                             the GPU for the mapped memory types; results checked on the host with closed forms
 * Exit code = number of failed checks.  Built and run by tests/test_zz_cpp_abi.py.
 */
#include <wholememory/embedding.h>
#include <wholememory/env_func_ptrs.h>
#include <wholememory/graph_op.h>
#include <wholememory/wholegraph_op.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
#include <wholememory/wholememory_tensor.h>

#include <cuda_runtime_api.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

static int g_failures = 0;
#define CHECK(cond)                                                                 \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++g_failures;                                                                 \
    }                                                                               \
  } while (0)
#define CHECK_OK(call) CHECK((call) == WHOLEMEMORY_SUCCESS)

/* ---- fork harness: rank bodies run in child processes, the unique id travels through a pipe (the reference's
 * MultiProcessRun + PipeBroadcast pattern, cpp/src/parallel_utils.cpp:46-100) ---- */
static int run_ranks(int world, const std::function<int(int, int, wholememory_unique_id_t)>& body)
{
  std::vector<std::pair<int, int>> pipes(world);
  for (auto& p : pipes) {
    int fds[2];
    if (pipe(fds) != 0) return 1;
    p = {fds[0], fds[1]};
  }
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {
    pid_t pid = fork();
    if (pid == 0) {
      wholememory_unique_id_t uid;
      memset(&uid, 0, sizeof(uid));
      if (r == 0) {
        if (wholememory_create_unique_id(&uid) != WHOLEMEMORY_SUCCESS) _exit(99);
        for (int q = 1; q < world; ++q)
          if (write(pipes[q].second, &uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(98);
      } else {
        if (read(pipes[r].first, &uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(97);
      }
      int rc = body(r, world, uid);
      fflush(stderr);
      _exit(rc > 90 ? 90 : rc);
    }
    kids.push_back(pid);
  }
  int failed = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
      fprintf(stderr, "rank process %d ended with status 0x%x\n", (int)k, st);
      failed += WIFEXITED(st) ? WEXITSTATUS(st) : 1;
    }
  }
  for (auto& p : pipes) {
    close(p.first);
    close(p.second);
  }
  return failed;
}

/* ---- caller-defined env functions (what a framework plugs in): cudaMalloc / cudaMallocHost / malloc ---- */
struct my_context {
  void* ptr = nullptr;
  wholememory_memory_allocation_type_t kind = WHOLEMEMORY_MA_NONE;
  wholememory_tensor_description_t desc;
};
static void my_create(void** ctx, void*) { *ctx = new my_context(); }
static void my_free(void* ctx, void*)
{
  auto* c = static_cast<my_context*>(ctx);
  if (c->ptr == nullptr) return;
  if (c->kind == WHOLEMEMORY_MA_DEVICE) cudaFree(c->ptr);
  else if (c->kind == WHOLEMEMORY_MA_PINNED) cudaFreeHost(c->ptr);
  else free(c->ptr);
  c->ptr = nullptr;
}
static void my_destroy(void* ctx, void* g)
{
  my_free(ctx, g);
  delete static_cast<my_context*>(ctx);
}
static void* my_malloc(wholememory_tensor_description_t* desc, wholememory_memory_allocation_type_t kind, void* ctx, void* g)
{
  auto* c = static_cast<my_context*>(ctx);
  my_free(ctx, g);
  c->desc      = *desc;
  c->kind      = kind;
  size_t bytes = (size_t)wholememory_get_memory_size_from_tensor(desc);
  if (bytes == 0) bytes = 1;
  if (kind == WHOLEMEMORY_MA_DEVICE) {
    if (cudaMalloc(&c->ptr, bytes) != cudaSuccess) c->ptr = nullptr;
  } else if (kind == WHOLEMEMORY_MA_PINNED) {
    if (cudaMallocHost(&c->ptr, bytes) != cudaSuccess) c->ptr = nullptr;
  } else {
    c->ptr = malloc(bytes);
  }
  return c->ptr;
}
static wholememory_env_func_t my_env = {{my_create, my_destroy, my_malloc, my_free, nullptr}, {my_malloc, my_free, nullptr}};

/* =============================================================== host-only section */
static int cpu_rank_body(int rank, int world, wholememory_unique_id_t uid)
{
  g_failures = 0;
  CHECK_OK(wholememory_init(0, LEVEL_ERROR));
  wholememory_comm_t comm = nullptr;
  CHECK_OK(wholememory_create_communicator(&comm, uid, rank, world));
  int r = -1, s = -1;
  CHECK_OK(wholememory_communicator_get_rank(&r, comm));
  CHECK_OK(wholememory_communicator_get_size(&s, comm));
  CHECK(r == rank && s == world);
  for (int i = 0; i < 50; ++i) CHECK_OK(wholememory_communicator_barrier(comm));
  CHECK(wholememory_is_intranode_communicator(comm));
  CHECK_OK(wholememory_destroy_communicator(comm));
  CHECK_OK(wholememory_finalize());
  return g_failures;
}

static void cpu_section()
{
  /* descriptors (reference tensor_description.cpp:20-233) */
  CHECK(wholememory_dtype_get_element_size(WHOLEMEMORY_DT_HALF) == 2 && wholememory_dtype_get_element_size(WHOLEMEMORY_DT_INT64) == 8);
  CHECK(wholememory_dtype_is_floating_number(WHOLEMEMORY_DT_BF16) && wholememory_dtype_is_integer_number(WHOLEMEMORY_DT_INT8));
  int64_t sz[2] = {100, 30};
  auto md       = wholememory_create_matrix_desc(sz, 32, 4, WHOLEMEMORY_DT_FLOAT);
  CHECK(wholememory_get_memory_element_count_from_matrix(&md) == 100 * 32);
  wholememory_tensor_description_t td;
  wholememory_copy_matrix_desc_to_tensor(&td, &md);
  CHECK(td.dim == 2 && td.sizes[1] == 30 && td.strides[0] == 32 && td.storage_offset == 4);
  wholememory_matrix_description_t back;
  CHECK(wholememory_convert_tensor_desc_to_matrix(&back, &td) && back.stride == 32);
  wholememory_array_description_t arr;
  CHECK(!wholememory_convert_tensor_desc_to_array(&arr, &td));
  CHECK(wholememory_unsqueeze_tensor(&td, 0) && td.dim == 3 && td.sizes[0] == 1);
  CHECK(wholememory_squeeze_tensor(&td, 0) && td.dim == 2);

  CHECK_OK(wholememory_init(0, LEVEL_ERROR));
  /* views over caller memory */
  std::vector<float> host(100 * 32, 0.f);
  wholememory_tensor_description_t hd;
  sz[0] = 100, sz[1] = 32;
  auto full = wholememory_create_matrix_desc(sz, 32, 0, WHOLEMEMORY_DT_FLOAT);
  wholememory_copy_matrix_desc_to_tensor(&hd, &full);
  wholememory_tensor_t t = nullptr, sub = nullptr;
  CHECK_OK(wholememory_make_tensor_from_pointer(&t, host.data(), &hd));
  CHECK(!wholememory_tensor_has_handle(t) && wholememory_tensor_get_data_pointer(t) == host.data());
  int64_t starts[2] = {10, 4}, ends[2] = {60, 20};
  CHECK_OK(wholememory_tensor_get_subtensor(t, starts, ends, &sub));
  auto* sd = wholememory_tensor_get_tensor_description(sub);
  CHECK(sd->sizes[0] == 50 && sd->sizes[1] == 16 && sd->strides[0] == 32 && sd->storage_offset == 10 * 32 + 4);
  CHECK(wholememory_tensor_get_root(sub) == t);
  CHECK_OK(wholememory_destroy_tensor(sub));

  /* env-function protocol with the built-in default env, HOST allocations */
  wholememory_env_func_t* env = wholememory::get_default_env_func();
  CHECK(env != nullptr && wholememory::get_cached_env_func() != nullptr);
  void* ctx = nullptr;
  env->temporary_fns.create_memory_context_fn(&ctx, env->temporary_fns.global_context);
  auto ad    = wholememory_create_array_desc(1000, 0, WHOLEMEMORY_DT_INT64);
  wholememory_tensor_description_t atd;
  wholememory_copy_array_desc_to_tensor(&atd, &ad);
  void* p = env->temporary_fns.malloc_fn(&atd, WHOLEMEMORY_MA_HOST, ctx, env->temporary_fns.global_context);
  CHECK(p != nullptr);
  if (p) memset(p, 0x5a, 8000);
  env->temporary_fns.free_fn(ctx, env->temporary_fns.global_context);
  env->temporary_fns.destroy_memory_context_fn(ctx, env->temporary_fns.global_context);

  /* single-rank communicator in this process; device ops must fail loudly when there is no GPU (no CPU fallback) */
  wholememory_unique_id_t uid;
  CHECK_OK(wholememory_create_unique_id(&uid));
  wholememory_comm_t comm = nullptr;
  CHECK_OK(wholememory_create_communicator(&comm, uid, 0, 1));
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 0;
  if (ndev == 0) {
    std::vector<int64_t> idx(8, 1);
    std::vector<float> out(8 * 32, 0.f);
    auto idesc = wholememory_create_array_desc(8, 0, WHOLEMEMORY_DT_INT64);
    wholememory_tensor_description_t itd, otd;
    wholememory_copy_array_desc_to_tensor(&itd, &idesc);
    int64_t osz[2] = {8, 32};
    auto om        = wholememory_create_matrix_desc(osz, 32, 0, WHOLEMEMORY_DT_FLOAT);
    wholememory_copy_matrix_desc_to_tensor(&otd, &om);
    wholememory_tensor_t it = nullptr, ot = nullptr;
    CHECK_OK(wholememory_make_tensor_from_pointer(&it, idx.data(), &itd));
    CHECK_OK(wholememory_make_tensor_from_pointer(&ot, out.data(), &otd));
    CHECK(wholememory_gather(t, it, ot, env, nullptr) != WHOLEMEMORY_SUCCESS);
    wholememory_tensor_t dev_t = nullptr;
    CHECK(wholememory_create_tensor(&dev_t, &hd, comm, WHOLEMEMORY_MT_CONTINUOUS, WHOLEMEMORY_ML_DEVICE) != WHOLEMEMORY_SUCCESS);
    CHECK_OK(wholememory_destroy_tensor(it));
    CHECK_OK(wholememory_destroy_tensor(ot));
  }
  CHECK_OK(wholememory_destroy_tensor(t));
  CHECK_OK(wholememory_destroy_communicator(comm));
  CHECK_OK(wholememory_finalize());

  /* two and three forked ranks: collective communicator creation, barriers, destruction */
  g_failures += run_ranks(2, cpu_rank_body);
  g_failures += run_ranks(3, cpu_rank_body);
}

/* =============================================================== GPU section */
static float pattern(int64_t row, int col) { return (float)((row * 7 + col) % 4093); } /* exact in fp32 */

template <typename T>
static T* to_device(const std::vector<T>& v)
{
  T* d = nullptr;
  cudaMalloc(reinterpret_cast<void**>(&d), std::max<size_t>(1, v.size() * sizeof(T)));
  cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

static wholememory_tensor_t wrap_device(void* p, int64_t rows, int64_t cols, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  if (cols > 0) {
    int64_t sz[2] = {rows, cols};
    auto m        = wholememory_create_matrix_desc(sz, cols, 0, dt);
    wholememory_copy_matrix_desc_to_tensor(&d, &m);
  } else {
    auto a = wholememory_create_array_desc(rows, 0, dt);
    wholememory_copy_array_desc_to_tensor(&d, &a);
  }
  wholememory_tensor_t t = nullptr;
  CHECK_OK(wholememory_make_tensor_from_pointer(&t, p, &d));
  return t;
}

static int gpu_rank_body(int rank, int world, wholememory_unique_id_t uid)
{
  g_failures = 0;
  CHECK(cudaSetDevice(0) == cudaSuccess);
  CHECK_OK(wholememory_init(0, LEVEL_ERROR));
  wholememory_comm_t comm = nullptr;
  CHECK_OK(wholememory_create_communicator(&comm, uid, rank, world));
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  const int64_t rows = 20011, cols = 96, n = 4099;
  std::mt19937_64 rng(0x5EED + rank);

  for (auto mt : {WHOLEMEMORY_MT_CONTINUOUS, WHOLEMEMORY_MT_CHUNKED, WHOLEMEMORY_MT_DISTRIBUTED}) {
    int64_t sz[2] = {rows, cols};
    auto md       = wholememory_create_matrix_desc(sz, cols, 0, WHOLEMEMORY_DT_FLOAT);
    wholememory_tensor_description_t td;
    wholememory_copy_matrix_desc_to_tensor(&td, &md);
    wholememory_tensor_t table = nullptr;
    CHECK_OK(wholememory_create_tensor(&table, &td, comm, mt, WHOLEMEMORY_ML_DEVICE));
    if (table == nullptr) continue;
    /* every rank fills its own shard */
    size_t first = 0, count = 0;
    CHECK_OK(wholememory_tensor_get_local_entry_start(&first, table));
    CHECK_OK(wholememory_tensor_get_local_entry_count(&count, table));
    void* local = nullptr;
    size_t local_bytes = 0, local_off = 0;
    CHECK_OK(wholememory_get_local_memory(&local, &local_bytes, &local_off, wholememory_tensor_get_memory_handle(table)));
    CHECK(local_bytes == count * cols * sizeof(float));
    std::vector<float> shard(count * cols);
    for (size_t i = 0; i < count; ++i)
      for (int c = 0; c < cols; ++c) shard[i * cols + c] = pattern((int64_t)(first + i), c);
    if (count > 0) cudaMemcpy(local, shard.data(), shard.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    CHECK_OK(wholememory_communicator_barrier(comm));

    /* gather: int64 indices over the WHOLE table, one negative (skipped), fp32 -> fp32 */
    std::vector<int64_t> idx(n);
    for (auto& v : idx) v = (int64_t)(rng() % rows);
    idx[5] = -1;
    std::vector<float> out_h(n * cols, -3.f);
    int64_t* d_idx = to_device(idx);
    float* d_out   = to_device(out_h);
    auto it = wrap_device(d_idx, n, 0, WHOLEMEMORY_DT_INT64);
    auto ot = wrap_device(d_out, n, cols, WHOLEMEMORY_DT_FLOAT);
    CHECK_OK(wholememory_gather(table, it, ot, wholememory::get_cached_env_func(), stream));
    cudaStreamSynchronize(stream);
    cudaMemcpy(out_h.data(), d_out, out_h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int64_t i = 0; i < n && ok; ++i)
      for (int c = 0; c < cols; ++c) {
        float want = idx[i] < 0 ? -3.f : pattern(idx[i], c);
        if (out_h[i * cols + c] != want) {
          ok = false;
          fprintf(stderr, "gather mismatch mt=%d row %ld col %d: %f vs %f\n", (int)mt, (long)i, c, out_h[i * cols + c], want);
          break;
        }
      }
    CHECK(ok);

    /* scatter (rank 0 only, distinct rows) then everybody gathers the rows back */
    CHECK_OK(wholememory_communicator_barrier(comm));
    const int64_t m = 512;
    std::vector<int64_t> sidx(m);
    for (int64_t i = 0; i < m; ++i) sidx[i] = (i * 37 + 11) % rows; /* 37 and rows are coprime: distinct */
    std::vector<float> src(m * cols);
    for (int64_t i = 0; i < m; ++i)
      for (int c = 0; c < cols; ++c) src[i * cols + c] = -(float)(i * 3 + c);
    int64_t* d_sidx = to_device(sidx);
    float* d_src    = to_device(src);
    auto st_i = wrap_device(d_sidx, m, 0, WHOLEMEMORY_DT_INT64);
    auto st_v = wrap_device(d_src, m, cols, WHOLEMEMORY_DT_FLOAT);
    if (rank == 0) {
      CHECK_OK(wholememory_scatter(st_v, st_i, table, wholememory::get_cached_env_func(), stream));
      cudaStreamSynchronize(stream);
    } else if (mt == WHOLEMEMORY_MT_DISTRIBUTED && world > 1) {
      /* DISTRIBUTED ops are collective in the ABI's contract (the exchange path needs every rank): take part with no rows */
      auto e_i = wrap_device(d_sidx, 0, 0, WHOLEMEMORY_DT_INT64);
      auto e_v = wrap_device(d_src, 0, cols, WHOLEMEMORY_DT_FLOAT);
      CHECK_OK(wholememory_scatter(e_v, e_i, table, wholememory::get_cached_env_func(), stream));
      cudaStreamSynchronize(stream);
      wholememory_destroy_tensor(e_i);
      wholememory_destroy_tensor(e_v);
    }
    CHECK_OK(wholememory_communicator_barrier(comm));
    std::vector<float> back(m * cols, 0.f);
    float* d_back = to_device(back);
    auto bt       = wrap_device(d_back, m, cols, WHOLEMEMORY_DT_FLOAT);
    CHECK_OK(wholememory_gather(table, st_i, bt, wholememory::get_cached_env_func(), stream));
    cudaStreamSynchronize(stream);
    cudaMemcpy(back.data(), d_back, back.size() * sizeof(float), cudaMemcpyDeviceToHost);
    CHECK(memcmp(back.data(), src.data(), back.size() * sizeof(float)) == 0);
    CHECK_OK(wholememory_communicator_barrier(comm));

    for (auto x : {it, ot, st_i, st_v, bt}) wholememory_destroy_tensor(x);
    cudaFree(d_idx), cudaFree(d_out), cudaFree(d_sidx), cudaFree(d_src), cudaFree(d_back);
    CHECK_OK(wholememory_destroy_tensor(table));
  }

  /* embedding + SGD step, values chosen so every product and sum is exact in fp32 (lr = 0.5, integer data) */
  {
    const int64_t erows = 3000, dim = 64, g = 1000;
    int64_t sz[2] = {erows, dim};
    auto md       = wholememory_create_matrix_desc(sz, dim, 0, WHOLEMEMORY_DT_FLOAT);
    wholememory_tensor_description_t td;
    wholememory_copy_matrix_desc_to_tensor(&td, &md);
    wholememory_embedding_t emb = nullptr;
    CHECK_OK(wholememory_create_embedding(&emb, &td, comm, WHOLEMEMORY_MT_CHUNKED, WHOLEMEMORY_ML_DEVICE, nullptr));
    wholememory_embedding_optimizer_t opt = nullptr;
    CHECK_OK(wholememory_create_embedding_optimizer(&opt, WHOLEMEMORY_OPT_SGD));
    CHECK_OK(wholememory_embedding_set_optimizer(emb, opt));
    wholememory_tensor_t wt = wholememory_embedding_get_embedding_tensor(emb);
    size_t first = 0, count = 0;
    CHECK_OK(wholememory_tensor_get_local_entry_start(&first, wt));
    CHECK_OK(wholememory_tensor_get_local_entry_count(&count, wt));
    void* local = nullptr;
    size_t lb = 0, lo = 0;
    CHECK_OK(wholememory_get_local_memory(&local, &lb, &lo, wholememory_tensor_get_memory_handle(wt)));
    const int64_t stride = wholememory_tensor_get_tensor_description(wt)->strides[0];
    std::vector<float> w(count * stride, 0.f);
    for (size_t i = 0; i < count; ++i)
      for (int c = 0; c < dim; ++c) w[i * stride + c] = (float)((first + i) % 100);
    if (count) cudaMemcpy(local, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    CHECK_OK(wholememory_communicator_barrier(comm));
    /* every rank sends the SAME g gradients (ids i*3 -> distinct, each appears `world` times across ranks) */
    std::vector<int64_t> gid(g);
    std::vector<float> grad(g * dim);
    for (int64_t i = 0; i < g; ++i) {
      gid[i] = (i * 3) % erows;
      for (int c = 0; c < dim; ++c) grad[i * dim + c] = (float)(2 * ((i + c) % 5));
    }
    int64_t* d_gid = to_device(gid);
    float* d_grad  = to_device(grad);
    auto gi = wrap_device(d_gid, g, 0, WHOLEMEMORY_DT_INT64);
    auto gv = wrap_device(d_grad, g, dim, WHOLEMEMORY_DT_FLOAT);
    CHECK_OK(wholememory_embedding_gather_gradient_apply(emb, gi, gv, false, 0.5f, &my_env, (int64_t)stream));
    cudaStreamSynchronize(stream);
    CHECK_OK(wholememory_communicator_barrier(comm));
    std::vector<float> got(g * dim, 0.f);
    float* d_got = to_device(got);
    auto go      = wrap_device(d_got, g, dim, WHOLEMEMORY_DT_FLOAT);
    CHECK_OK(wholememory_embedding_gather(emb, gi, go, false, &my_env, (int64_t)stream));
    cudaStreamSynchronize(stream);
    cudaMemcpy(got.data(), d_got, got.size() * sizeof(float), cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int64_t i = 0; i < g && ok; ++i)
      for (int c = 0; c < dim; ++c) {
        /* w - lr * (sum over ranks of the same gradient) = w - 0.5 * world * grad */
        float want = (float)(gid[i] % 100) - 0.5f * (float)world * grad[i * dim + c];
        if (got[i * dim + c] != want) {
          ok = false;
          fprintf(stderr, "SGD mismatch row %ld col %d: %f vs %f\n", (long)gid[i], c, got[i * dim + c], want);
          break;
        }
      }
    CHECK(ok);
    CHECK_OK(wholememory_communicator_barrier(comm));
    for (auto x : {gi, gv, go}) wholememory_destroy_tensor(x);
    cudaFree(d_gid), cudaFree(d_grad), cudaFree(d_got);
    CHECK_OK(wholememory_destroy_embedding(emb));
    wholememory_destroy_embedding_optimizer(opt);
  }

  /* CSR neighbor sampling with caller-defined output contexts: structural invariants (offsets = scan of min(deg, k),
   * samples are distinct positions of the center's own neighbour list, deg <= k copies the list in order) */
  {
    const int64_t nodes = 2000, centers_n = 500;
    const int k = 10;
    std::vector<int64_t> row_ptr(nodes + 1, 0);
    for (int64_t v = 0; v < nodes; ++v) row_ptr[v + 1] = row_ptr[v] + (int64_t)((v * 2654435761u) % 40);
    const int64_t edges = row_ptr[nodes];
    std::vector<int64_t> col(edges);
    for (int64_t e = 0; e < edges; ++e) col[e] = e; /* neighbour "id" = its own edge index: positions are recoverable */
    auto rp_d = wholememory_create_array_desc(nodes + 1, 0, WHOLEMEMORY_DT_INT64);
    auto cp_d = wholememory_create_array_desc(edges, 0, WHOLEMEMORY_DT_INT64);
    wholememory_tensor_description_t rp_td, cp_td;
    wholememory_copy_array_desc_to_tensor(&rp_td, &rp_d);
    wholememory_copy_array_desc_to_tensor(&cp_td, &cp_d);
    wholememory_tensor_t rp = nullptr, cp = nullptr;
    CHECK_OK(wholememory_create_tensor(&rp, &rp_td, comm, WHOLEMEMORY_MT_CHUNKED, WHOLEMEMORY_ML_DEVICE));
    CHECK_OK(wholememory_create_tensor(&cp, &cp_td, comm, WHOLEMEMORY_MT_CHUNKED, WHOLEMEMORY_ML_DEVICE));
    for (auto pr : {std::make_pair(rp, (const void*)row_ptr.data()), std::make_pair(cp, (const void*)col.data())}) {
      size_t first = 0, count = 0;
      CHECK_OK(wholememory_tensor_get_local_entry_start(&first, pr.first));
      CHECK_OK(wholememory_tensor_get_local_entry_count(&count, pr.first));
      void* local = nullptr;
      size_t lb = 0, lo = 0;
      CHECK_OK(wholememory_get_local_memory(&local, &lb, &lo, wholememory_tensor_get_memory_handle(pr.first)));
      if (count) cudaMemcpy(local, static_cast<const int64_t*>(pr.second) + first, count * sizeof(int64_t), cudaMemcpyHostToDevice);
    }
    cudaDeviceSynchronize();
    CHECK_OK(wholememory_communicator_barrier(comm));
    std::vector<int64_t> centers(centers_n);
    for (auto& c : centers) c = (int64_t)(rng() % nodes);
    int64_t* d_c = to_device(centers);
    std::vector<int> offs(centers_n + 1, -1);
    int* d_offs = to_device(offs);
    auto ct = wrap_device(d_c, centers_n, 0, WHOLEMEMORY_DT_INT64);
    auto ot = wrap_device(d_offs, centers_n + 1, 0, WHOLEMEMORY_DT_INT);
    my_context dst_ctx, lid_ctx, gid_ctx;
    CHECK_OK(wholegraph_csr_unweighted_sample_without_replacement(rp, cp, ct, k, ot, &dst_ctx, &lid_ctx, &gid_ctx, 1234ull + rank,
                                                                  &my_env, stream));
    cudaStreamSynchronize(stream);
    cudaMemcpy(offs.data(), d_offs, offs.size() * sizeof(int), cudaMemcpyDeviceToHost);
    const int total = offs[centers_n];
    CHECK(offs[0] == 0 && total >= 0 && dst_ctx.desc.sizes[0] == total && dst_ctx.desc.dtype == WHOLEMEMORY_DT_INT64);
    CHECK(lid_ctx.desc.dtype == WHOLEMEMORY_DT_INT && gid_ctx.desc.dtype == WHOLEMEMORY_DT_INT64);
    std::vector<int64_t> dst(std::max(total, 1)), egid(std::max(total, 1));
    std::vector<int> lid(std::max(total, 1));
    if (total > 0) {
      cudaMemcpy(dst.data(), dst_ctx.ptr, total * sizeof(int64_t), cudaMemcpyDeviceToHost);
      cudaMemcpy(egid.data(), gid_ctx.ptr, total * sizeof(int64_t), cudaMemcpyDeviceToHost);
      cudaMemcpy(lid.data(), lid_ctx.ptr, total * sizeof(int), cudaMemcpyDeviceToHost);
    }
    bool ok = true;
    for (int64_t i = 0; i < centers_n && ok; ++i) {
      const int64_t b = row_ptr[centers[i]], deg = row_ptr[centers[i] + 1] - b;
      const int want = (int)std::min<int64_t>(deg, k);
      if (offs[i + 1] - offs[i] != want) ok = false;
      std::vector<char> seen(deg > 0 ? deg : 1, 0);
      for (int j = offs[i]; j < offs[i + 1] && ok; ++j) {
        const int64_t pos = dst[j] - b; /* col[e] == e */
        if (pos < 0 || pos >= deg || seen[pos] || egid[j] != dst[j] || lid[j] != (int)i) ok = false;
        else seen[pos] = 1;
        if (deg <= k && pos != j - offs[i]) ok = false; /* whole list, in order */
      }
    }
    CHECK(ok);
    my_free(&dst_ctx, nullptr), my_free(&lid_ctx, nullptr), my_free(&gid_ctx, nullptr);
    CHECK_OK(wholememory_communicator_barrier(comm));
    for (auto x : {ct, ot}) wholememory_destroy_tensor(x);
    cudaFree(d_c), cudaFree(d_offs);
    CHECK_OK(wholememory_destroy_tensor(rp));
    CHECK_OK(wholememory_destroy_tensor(cp));
  }

  wholememory::drop_cached_env_func_cache();
  cudaStreamDestroy(stream);
  CHECK_OK(wholememory_destroy_communicator(comm));
  CHECK_OK(wholememory_finalize());
  return g_failures;
}

int main(int argc, char** argv)
{
  const char* mode = argc > 1 ? argv[1] : "cpu";
  if (strcmp(mode, "cpu") == 0) {
    cpu_section();
  } else if (strcmp(mode, "gpu") == 0) {
    int ranks = argc > 2 ? atoi(argv[2]) : 1;
    g_failures += run_ranks(ranks, gpu_rank_body); /* fork BEFORE any CUDA call in this process */
  } else {
    fprintf(stderr, "usage: %s cpu | gpu [ranks]\n", argv[0]);
    return 2;
  }
  printf("abi_cpp_test %s: %d failed checks\n", mode, g_failures);
  return g_failures > 100 ? 100 : g_failures;
}
