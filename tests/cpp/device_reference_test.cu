/*
 * wholememory::device_reference<T> (include/wholememory/device_reference.cuh, the public device helper downstream kernels
 * use; reference cpp/include/wholememory/device_reference.cuh:25-71) exercised IN A KERNEL on the grefs this library
 * hands out: CONTINUOUS (flat pointer), CHUNKED with an equal partition (same_chunk: owner = offset / stride) and CHUNKED
 * with a custom partition (offset-table lookup), at 1 and 3 ranks sharing the GPU.
 * Every rank writes element i = f(i) through ref[i] for the elements it owns (by index range), all ranks then read the
 * WHOLE table through ref[i] and count mismatches; the result is also compared with what the local-memory accessor sees.
 *   device_reference_test <ranks>      exit code = number of failed checks
 */
#include <wholememory/device_reference.cuh>
#include <wholememory/wholememory.h>
#include <wholememory/wholememory_tensor.h>

#include <cuda_runtime_api.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
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

__host__ __device__ inline float value_of(size_t i) { return (float)(i % 1000003) * 0.5f + 1.0f; }

__global__ void write_kernel(wholememory_gref_t g, size_t begin, size_t end)
{
  wholememory::device_reference<float> ref(g);
  for (size_t i = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x) ref[i] = value_of(i);
}

__global__ void check_kernel(wholememory_gref_t g, size_t n, unsigned long long* bad)
{
  wholememory::device_reference<float> ref(g);
  unsigned long long mine = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) mine += ref[i] != value_of(i);
  if (mine) atomicAdd(bad, mine);
}

static int rank_body(int rank, int world, wholememory_unique_id_t uid)
{
  if (cudaSetDevice(0) != cudaSuccess) return 50;
  CHECK_OK(wholememory_init(0, LEVEL_WARN));
  wholememory_comm_t comm = nullptr;
  CHECK_OK(wholememory_create_communicator(&comm, uid, rank, world));
  if (comm == nullptr) return 51;
  const size_t rows = 30011, cols = 32, n = rows * cols; /* rows chosen so the equal split leaves a short tail rank */
  std::vector<size_t> custom(world);
  {
    size_t left = rows;
    for (int r = 0; r < world; ++r) {
      custom[r] = r + 1 == world ? left : (rows / world / 2) * (size_t)(r + 1) + 7;
      left -= custom[r];
    }
  }
  struct variant {
    wholememory_memory_type_t type;
    bool use_custom;
    const char* name;
  } variants[] = {{WHOLEMEMORY_MT_CONTINUOUS, false, "continuous"},
                  {WHOLEMEMORY_MT_CHUNKED, false, "chunked, equal partition"},
                  {WHOLEMEMORY_MT_CHUNKED, true, "chunked, custom partition"}};
  unsigned long long* d_bad = nullptr;
  if (cudaMalloc(&d_bad, sizeof(*d_bad)) != cudaSuccess) return 52;
  for (auto& v : variants) {
    if (v.use_custom && world == 1) continue;
    wholememory_tensor_description_t d;
    wholememory_initialize_tensor_desc(&d);
    d.dim = 2, d.dtype = WHOLEMEMORY_DT_FLOAT, d.sizes[0] = (int64_t)rows, d.sizes[1] = (int64_t)cols, d.strides[0] = (int64_t)cols, d.strides[1] = 1;
    wholememory_tensor_t t = nullptr;
    CHECK_OK(wholememory_create_tensor(&t, &d, comm, v.type, WHOLEMEMORY_ML_DEVICE, v.use_custom ? custom.data() : nullptr));
    if (t == nullptr) return 53;
    wholememory_gref_t g;
    CHECK_OK(wholememory_tensor_get_global_reference(t, &g));
    if (v.type == WHOLEMEMORY_MT_CONTINUOUS) CHECK(g.stride == 0);
    else {
      CHECK(g.stride > 0 && g.world_size == world);
      if (world > 1) CHECK(g.same_chunk == !v.use_custom);
    }
    /* each rank writes the index range [rank, rank+1) / world of the WHOLE table through the reference: most of it
     * lands in other ranks' memory */
    size_t begin = n * (size_t)rank / world, end = n * (size_t)(rank + 1) / world;
    write_kernel<<<148, 256>>>(g, begin, end);
    CHECK(cudaDeviceSynchronize() == cudaSuccess);
    CHECK_OK(wholememory_communicator_barrier(comm));
    CHECK(cudaMemset(d_bad, 0, sizeof(*d_bad)) == cudaSuccess);
    check_kernel<<<148, 256>>>(g, n, d_bad);
    unsigned long long bad = ~0ull;
    CHECK(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost) == cudaSuccess);
    if (bad != 0) fprintf(stderr, "rank %d, %s: %llu elements differ through device_reference\n", rank, v.name, bad);
    CHECK(bad == 0);
    /* the same bytes through the library's own accessor: my shard starts at my first row */
    size_t start = 0, count = 0;
    CHECK_OK(wholememory_tensor_get_local_entry_start(&start, t));
    CHECK_OK(wholememory_tensor_get_local_entry_count(&count, t));
    if (v.use_custom) CHECK(count == custom[rank]);
    wholememory_tensor_t local = nullptr;
    CHECK_OK(wholememory_tensor_map_local_tensor(t, &local));
    if (count > 0 && local != nullptr) {
      std::vector<float> host(count * cols);
      CHECK(cudaMemcpy(host.data(), wholememory_tensor_get_data_pointer(local), host.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess);
      size_t wrong = 0;
      for (size_t k = 0; k < host.size(); ++k) wrong += host[k] != value_of(start * cols + k);
      CHECK(wrong == 0);
    }
    if (local) CHECK_OK(wholememory_destroy_tensor(local));
    CHECK_OK(wholememory_communicator_barrier(comm));
    CHECK_OK(wholememory_destroy_tensor(t));
  }
  cudaFree(d_bad);
  CHECK_OK(wholememory_destroy_communicator(comm));
  return g_failures;
}

int main(int argc, char** argv)
{
  int world = argc > 1 ? atoi(argv[1]) : 1;
  /* fork BEFORE any CUDA call in this process; the unique id travels through pipes */
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
      } else if (read(pipes[r].first, &uid, sizeof(uid)) != (ssize_t)sizeof(uid)) {
        _exit(97);
      }
      int rc = rank_body(r, world, uid);
      fflush(stderr);
      _exit(rc > 90 ? 90 : rc);
    }
    kids.push_back(pid);
  }
  int failed = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) failed += WIFEXITED(st) ? WEXITSTATUS(st) : 1;
  }
  printf("%d failed checks\n", failed);
  return failed;
}
