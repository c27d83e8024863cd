/*
 * C++ gather / scatter benchmark with the command line of the reference's GATHER_SCATTER_BENCH
 * (cpp/bench/wholememory_ops/gather_scatter_bench.cu:420-470: -t memory type, -l location, -e table bytes, -g gather
 * bytes, -d embedding dim, -c loop count, -f gather|scatter, -n GPUs, -m partition method) for ONE node, plus
 *   --lib PATH   the shared library to drive.  It is dlopen'ed and only extern "C" symbols of the C ABI are looked up, so
 *                the SAME binary times this repo's wholegraph_b200/lib/libwholegraph.so and the reference's own build
 *                (oracle/_ref/libwholegraph_ref.so) -- the "same harness, same indices" comparison of BASELINE.md section 3.
 * One forked process per GPU (unique id over pipes, like the reference's MultiProcessRun); fp32 table filled with a
 * row-derived pattern; int64 indices from mt19937_64(0x5EED + rank), uniform over the whole table.
 * Reported per rank and as min / max / avg:  the reference's number, "Bandwidth" = gathered bytes / host wall time over
 * loop_count back-to-back calls + one device sync (gather_scatter_bench.cu:363-366), and the same bytes / CUDA-event time.
 *
 * Build:  g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include tools/gather_scatter_bench.cpp -o gather_scatter_bench \
 *             -L/usr/local/cuda/lib64 -lcudart -ldl
 * Example (config C2):  ./gather_scatter_bench -t 1 -l 1 -e 102400000000 -g 1073741824 -d 256 -c 20 -n 1
 */
#include <wholememory/env_func_ptrs.h>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_op.h>
#include <wholememory/wholememory_tensor.h>

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <getopt.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

struct params_t {
  int memory_type       = 2; /* CHUNKED  (reference default, gather_scatter_bench.cu:206-229) */
  int memory_location   = 1; /* DEVICE */
  int64_t table_bytes   = 1024000;
  int64_t gather_bytes  = 1024;
  int64_t dim           = 32;
  int loop_count        = 20;
  std::string test_type = "gather";
  int num_gpu           = 0; /* 0 = all visible */
  int partition_method  = 0;
  std::string lib       = "wholegraph_b200/lib/libwholegraph.so";
};

/* the C ABI, resolved by name from whichever library was given */
struct abi_t {
  void* so = nullptr;
  decltype(&wholememory_init) init;
  decltype(&wholememory_finalize) finalize;
  decltype(&wholememory_create_unique_id) create_unique_id;
  decltype(&wholememory_create_communicator) create_communicator;
  decltype(&wholememory_destroy_communicator) destroy_communicator;
  decltype(&wholememory_communicator_barrier) barrier;
  decltype(&wholememory_create_tensor) create_tensor;
  decltype(&wholememory_destroy_tensor) destroy_tensor;
  decltype(&wholememory_make_tensor_from_pointer) make_tensor_from_pointer;
  decltype(&wholememory_tensor_get_memory_handle) tensor_get_memory_handle;
  decltype(&wholememory_tensor_get_local_entry_start) local_entry_start;
  decltype(&wholememory_tensor_get_local_entry_count) local_entry_count;
  decltype(&wholememory_get_local_memory) get_local_memory;
  decltype(&wholememory_gather) gather;
  decltype(&wholememory_scatter) scatter;
};

template <typename F>
bool sym(void* so, const char* name, F* out)
{
  *out = reinterpret_cast<F>(dlsym(so, name));
  if (*out == nullptr) fprintf(stderr, "symbol %s not found: %s\n", name, dlerror());
  return *out != nullptr;
}

bool load_abi(const std::string& path, abi_t* a)
{
  a->so = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
  if (a->so == nullptr) {
    fprintf(stderr, "cannot load %s: %s\n", path.c_str(), dlerror());
    return false;
  }
  return sym(a->so, "wholememory_init", &a->init) && sym(a->so, "wholememory_finalize", &a->finalize) &&
         sym(a->so, "wholememory_create_unique_id", &a->create_unique_id) &&
         sym(a->so, "wholememory_create_communicator", &a->create_communicator) &&
         sym(a->so, "wholememory_destroy_communicator", &a->destroy_communicator) &&
         sym(a->so, "wholememory_communicator_barrier", &a->barrier) && sym(a->so, "wholememory_create_tensor", &a->create_tensor) &&
         sym(a->so, "wholememory_destroy_tensor", &a->destroy_tensor) &&
         sym(a->so, "wholememory_make_tensor_from_pointer", &a->make_tensor_from_pointer) &&
         sym(a->so, "wholememory_tensor_get_memory_handle", &a->tensor_get_memory_handle) &&
         sym(a->so, "wholememory_tensor_get_local_entry_start", &a->local_entry_start) &&
         sym(a->so, "wholememory_tensor_get_local_entry_count", &a->local_entry_count) &&
         sym(a->so, "wholememory_get_local_memory", &a->get_local_memory) && sym(a->so, "wholememory_gather", &a->gather) &&
         sym(a->so, "wholememory_scatter", &a->scatter);
}

/* descriptors are plain structs: filled here so that nothing but the entry points above comes from the library */
wholememory_tensor_description_t matrix_desc(int64_t rows, int64_t cols, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  memset(&d, 0, sizeof(d));
  d.dim = 2, d.dtype = dt, d.storage_offset = 0;
  d.sizes[0] = rows, d.sizes[1] = cols, d.strides[0] = cols, d.strides[1] = 1;
  return d;
}
wholememory_tensor_description_t array_desc(int64_t n, wholememory_dtype_t dt)
{
  wholememory_tensor_description_t d;
  memset(&d, 0, sizeof(d));
  d.dim = 1, d.dtype = dt, d.storage_offset = 0;
  d.sizes[0] = n, d.strides[0] = 1;
  return d;
}

/* caller-side env functions: cudaMalloc / cudaMallocHost / malloc with a private context */
struct ctx_t {
  void* p = nullptr;
  int kind = 0;
};
void env_create(void** c, void*) { *c = new ctx_t(); }
void env_free(void* c, void*)
{
  auto* x = static_cast<ctx_t*>(c);
  if (x->p == nullptr) return;
  if (x->kind == WHOLEMEMORY_MA_DEVICE) cudaFree(x->p);
  else if (x->kind == WHOLEMEMORY_MA_PINNED) cudaFreeHost(x->p);
  else free(x->p);
  x->p = nullptr;
}
void env_destroy(void* c, void* g)
{
  env_free(c, g);
  delete static_cast<ctx_t*>(c);
}
void* env_malloc(wholememory_tensor_description_t* d, wholememory_memory_allocation_type_t kind, void* c, void* g)
{
  auto* x = static_cast<ctx_t*>(c);
  env_free(c, g);
  size_t elems = 1;
  for (int i = 0; i < d->dim; ++i) elems *= (size_t)d->sizes[i];
  static const size_t esize[] = {0, 4, 2, 8, 2, 4, 8, 2, 1};
  size_t bytes = std::max<size_t>(1, elems * esize[d->dtype]);
  x->kind      = kind;
  if (kind == WHOLEMEMORY_MA_DEVICE) {
    if (cudaMalloc(&x->p, bytes) != cudaSuccess) x->p = nullptr;
  } else if (kind == WHOLEMEMORY_MA_PINNED) {
    if (cudaMallocHost(&x->p, bytes) != cudaSuccess) x->p = nullptr;
  } else {
    x->p = malloc(bytes);
  }
  return x->p;
}
wholememory_env_func_t g_env = {{env_create, env_destroy, env_malloc, env_free, nullptr}, {env_malloc, env_free, nullptr}};

double now_us()
{
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec * 1e6 + tv.tv_usec;
}

#define OK(call)                                                                     \
  do {                                                                               \
    if ((call) != WHOLEMEMORY_SUCCESS) {                                             \
      fprintf(stderr, "rank %d: %s failed (%s:%d)\n", rank, #call, __FILE__, __LINE__); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int rank_main(const params_t& p, int rank, int world, wholememory_unique_id_t uid, int result_fd)
{
  abi_t a;
  if (!load_abi(p.lib, &a)) return 1;
  if (cudaSetDevice(rank) != cudaSuccess) return 1;
  OK(a.init(0, LEVEL_WARN));
  wholememory_comm_t comm = nullptr;
  OK(a.create_communicator(&comm, uid, rank, world));
  const int64_t row_bytes = p.dim * 4;
  const int64_t rows      = (p.table_bytes + row_bytes - 1) / row_bytes;
  const int64_t n         = std::max<int64_t>(1, (p.gather_bytes + row_bytes - 1) / row_bytes);
  auto td                 = matrix_desc(rows, p.dim, WHOLEMEMORY_DT_FLOAT);
  std::vector<size_t> part;
  if (p.partition_method == 1 && world > 1) { /* random partition: every rank derives the same split */
    std::mt19937_64 g(1234);
    part.assign(world, 1);
    size_t left = (size_t)rows - world;
    for (int r = 0; r < world - 1; ++r) {
      size_t take = left ? g() % (2 * left / (world - r) + 1) : 0;
      take        = std::min(take, left);
      part[r] += take;
      left -= take;
    }
    part[world - 1] += left;
  }
  wholememory_tensor_t table = nullptr;
  OK(a.create_tensor(&table, &td, comm, (wholememory_memory_type_t)p.memory_type, (wholememory_memory_location_t)p.memory_location,
                     part.empty() ? nullptr : part.data()));
  size_t first = 0, count = 0;
  OK(a.local_entry_start(&first, table));
  OK(a.local_entry_count(&count, table));
  void* local = nullptr;
  size_t lb = 0, lo = 0;
  OK(a.get_local_memory(&local, &lb, &lo, a.tensor_get_memory_handle(table)));
  { /* fill my shard, 64 MiB at a time: element (r, c) = r & 0xffffff */
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / row_bytes);
    std::vector<float> buf(chunk_rows * p.dim);
    for (size_t s = 0; s < count; s += chunk_rows) {
      size_t m = std::min(chunk_rows, count - s);
      for (size_t i = 0; i < m; ++i) std::fill_n(buf.begin() + i * p.dim, p.dim, (float)((first + s + i) & 0xffffff));
      if (p.memory_location == 1) cudaMemcpy(static_cast<char*>(local) + s * row_bytes, buf.data(), m * row_bytes, cudaMemcpyHostToDevice);
      else memcpy(static_cast<char*>(local) + s * row_bytes, buf.data(), m * row_bytes);
    }
  }
  cudaDeviceSynchronize();
  OK(a.barrier(comm));

  std::mt19937_64 g(0x5EED + rank);
  std::vector<int64_t> idx(n);
  for (auto& v : idx) v = (int64_t)(g() % (uint64_t)rows);
  const bool scatter = p.test_type == "scatter";
  if (scatter) { /* distinct rows per rank and across ranks: stride walk from a rank-specific start */
    for (int64_t i = 0; i < n; ++i) idx[i] = (int64_t)(((uint64_t)i * world + rank) % (uint64_t)rows);
  }
  int64_t* d_idx = nullptr;
  float* d_rows  = nullptr;
  cudaMalloc(reinterpret_cast<void**>(&d_idx), n * sizeof(int64_t));
  cudaMalloc(reinterpret_cast<void**>(&d_rows), n * row_bytes);
  cudaMemcpy(d_idx, idx.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice);
  cudaMemset(d_rows, 0, n * row_bytes);
  auto id = array_desc(n, WHOLEMEMORY_DT_INT64);
  auto od = matrix_desc(n, p.dim, WHOLEMEMORY_DT_FLOAT);
  wholememory_tensor_t it = nullptr, ot = nullptr;
  OK(a.make_tensor_from_pointer(&it, d_idx, &id));
  OK(a.make_tensor_from_pointer(&ot, d_rows, &od));
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  auto call = [&]() { return scatter ? a.scatter(ot, it, table, &g_env, stream, -1) : a.gather(table, it, ot, &g_env, stream, -1); };
  for (int w = 0; w < 5; ++w) OK(call());
  cudaStreamSynchronize(stream);
  if (!scatter) { /* spot check against the pattern */
    std::vector<float> chk(p.dim);
    cudaMemcpy(chk.data(), d_rows + (n - 1) * p.dim, row_bytes, cudaMemcpyDeviceToHost);
    if (chk[0] != (float)(idx[n - 1] & 0xffffff) || chk[p.dim - 1] != chk[0]) {
      fprintf(stderr, "rank %d: gathered row does not match the pattern\n", rank);
      return 1;
    }
  }
  OK(a.barrier(comm));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double t0 = now_us();
  cudaEventRecord(e0, stream);
  for (int l = 0; l < p.loop_count; ++l) OK(call());
  cudaEventRecord(e1, stream);
  cudaDeviceSynchronize();
  double host_us = (now_us() - t0) / p.loop_count;
  float dev_ms   = 0.f;
  cudaEventElapsedTime(&dev_ms, e0, e1);
  double dev_us  = dev_ms * 1e3 / p.loop_count;
  OK(a.barrier(comm));
  double res[2] = {n * (double)row_bytes / host_us / 1e3, n * (double)row_bytes / dev_us / 1e3}; /* GB/s */
  if (write(result_fd, res, sizeof(res)) != (ssize_t)sizeof(res)) return 1;

  a.destroy_tensor(it);
  a.destroy_tensor(ot);
  cudaFree(d_idx);
  cudaFree(d_rows);
  OK(a.destroy_tensor(table));
  OK(a.destroy_communicator(comm));
  a.finalize();
  return 0;
}

void usage(const char* argv0)
{
  printf("Usage: %s [options]\n"
         "  -h, --help\n"
         "  -t, --memory_type           1: Continuous, 2: Chunked, 3: Distributed (default 2)\n"
         "  -l, --memory_location       1: Device, 2: Host (default 1)\n"
         "  -e, --embedding_table_size  table bytes (default 1024000)\n"
         "  -g, --gather_size           gathered bytes per call and rank (default 1024)\n"
         "  -d, --embedding_dim         fp32 elements per row (default 32)\n"
         "  -c, --loop_count            timed calls (default 20)\n"
         "  -f, --test_type             gather | scatter\n"
         "  -n, --num_gpu               ranks = GPUs used (default: all visible)\n"
         "  -m, --partition_method      0: equal, 1: random\n"
         "      --lib PATH              libwholegraph.so to drive (this repo's by default; oracle/_ref/libwholegraph_ref.so = the reference)\n"
         "  (-r/-s/-a/-p/-b of the reference bench are accepted and ignored: one node, NCCL-free control plane)\n",
         argv0);
}

}  // namespace

int main(int argc, char** argv)
{
  params_t p;
  const option opts[] = {{"help", no_argument, nullptr, 'h'},
                         {"memory_type", required_argument, nullptr, 't'},
                         {"memory_location", required_argument, nullptr, 'l'},
                         {"embedding_table_size", required_argument, nullptr, 'e'},
                         {"gather_size", required_argument, nullptr, 'g'},
                         {"embedding_dim", required_argument, nullptr, 'd'},
                         {"loop_count", required_argument, nullptr, 'c'},
                         {"test_type", required_argument, nullptr, 'f'},
                         {"node_rank", required_argument, nullptr, 'r'},
                         {"node_size", required_argument, nullptr, 's'},
                         {"num_gpu", required_argument, nullptr, 'n'},
                         {"server_addr", required_argument, nullptr, 'a'},
                         {"server_port", required_argument, nullptr, 'p'},
                         {"partition_method", required_argument, nullptr, 'm'},
                         {"distributed_backend", required_argument, nullptr, 'b'},
                         {"lib", required_argument, nullptr, 'L'},
                         {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "ht:l:e:g:d:c:f:a:p:r:s:n:m:b:", opts, nullptr)) != -1) {
    switch (c) {
      case 'h': usage(argv[0]); return 0;
      case 't': p.memory_type = atoi(optarg); break;
      case 'l': p.memory_location = atoi(optarg); break;
      case 'e': p.table_bytes = atoll(optarg); break;
      case 'g': p.gather_bytes = atoll(optarg); break;
      case 'd': p.dim = atoll(optarg); break;
      case 'c': p.loop_count = atoi(optarg); break;
      case 'f': p.test_type = optarg; break;
      case 'n': p.num_gpu = atoi(optarg); break;
      case 'm': p.partition_method = atoi(optarg); break;
      case 'L': p.lib = optarg; break;
      case 'r':
      case 's':
      case 'a':
      case 'p':
      case 'b': break;
      default: usage(argv[0]); return 2;
    }
  }
  if (p.memory_type < 1 || p.memory_type > 3 || p.memory_location < 1 || p.memory_location > 2 || p.dim <= 0 || p.loop_count <= 0 ||
      (p.test_type != "gather" && p.test_type != "scatter")) {
    usage(argv[0]);
    return 2;
  }
  int world = p.num_gpu;
  if (world <= 0) { /* count devices in a throw-away child so this process never creates a CUDA context (reference ForkGetDeviceCount) */
    int fds[2];
    if (pipe(fds) != 0) return 1;
    pid_t pid = fork();
    if (pid == 0) {
      int n = 0;
      if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
      if (write(fds[1], &n, sizeof(n)) != (ssize_t)sizeof(n)) _exit(1);
      _exit(0);
    }
    if (read(fds[0], &world, sizeof(world)) != (ssize_t)sizeof(world)) world = 0;
    waitpid(pid, nullptr, 0);
    if (world <= 0) {
      fprintf(stderr, "no CUDA device visible\n");
      return 1;
    }
  }
  std::vector<int> uid_r(world), uid_w(world), res_r(world), res_w(world);
  for (int r = 0; r < world; ++r) {
    int a[2], b[2];
    if (pipe(a) != 0 || pipe(b) != 0) return 1;
    uid_r[r] = a[0], uid_w[r] = a[1], res_r[r] = b[0], res_w[r] = b[1];
  }
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {
    pid_t pid = fork();
    if (pid == 0) {
      wholememory_unique_id_t uid;
      memset(&uid, 0, sizeof(uid));
      if (r == 0) {
        abi_t a;
        if (!load_abi(p.lib, &a) || a.create_unique_id(&uid) != WHOLEMEMORY_SUCCESS) _exit(1);
        for (int q = 1; q < world; ++q)
          if (write(uid_w[q], &uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(1);
      } else if (read(uid_r[r], &uid, sizeof(uid)) != (ssize_t)sizeof(uid)) {
        _exit(1);
      }
      _exit(rank_main(p, r, world, uid, res_w[r]));
    }
    kids.push_back(pid);
  }
  int bad = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad;
  }
  if (bad) {
    fprintf(stderr, "%d rank process(es) failed\n", bad);
    return 1;
  }
  double host_min = 1e30, host_max = 0, host_sum = 0, dev_min = 1e30, dev_max = 0, dev_sum = 0;
  for (int r = 0; r < world; ++r) {
    double res[2];
    if (read(res_r[r], res, sizeof(res)) != (ssize_t)sizeof(res)) return 1;
    printf("rank %d: Bandwidth (host-timed, reference definition) %.2f GB/s, device-timed %.2f GB/s\n", r, res[0], res[1]);
    host_min = std::min(host_min, res[0]), host_max = std::max(host_max, res[0]), host_sum += res[0];
    dev_min = std::min(dev_min, res[1]), dev_max = std::max(dev_max, res[1]), dev_sum += res[1];
  }
  printf("%s, %s, type %d location %d, table %ld B, %ld B per call, dim %ld, %d loops, %d GPU(s)\n", p.lib.c_str(), p.test_type.c_str(),
         p.memory_type, p.memory_location, (long)p.table_bytes, (long)p.gather_bytes, (long)p.dim, p.loop_count, world);
  printf("host-timed   GB/s per GPU: min %.2f max %.2f avg %.2f   aggregate %.2f\n", host_min, host_max, host_sum / world, host_sum);
  printf("device-timed GB/s per GPU: min %.2f max %.2f avg %.2f   aggregate %.2f\n", dev_min, dev_max, dev_sum / world, dev_sum);
  return 0;
}
