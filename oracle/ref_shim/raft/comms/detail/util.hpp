#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <raft/core/comms.hpp>
#include <raft/core/error.hpp>
#include <string>

#define RAFT_NCCL_TRY(call)                                                                    \
  do {                                                                                         \
    ncclResult_t const raft_nccl_status_ = (call);                                             \
    if (raft_nccl_status_ != ncclSuccess) {                                                    \
      throw raft::logic_error(std::string("NCCL error: ") + ncclGetErrorString(raft_nccl_status_)); \
    }                                                                                          \
  } while (0)

namespace raft {
namespace comms {
namespace detail {
inline status_t nccl_sync_stream(ncclComm_t comm, cudaStream_t stream)
{
  while (true) {
    cudaError_t q = cudaStreamQuery(stream);
    if (q == cudaSuccess) return status_t::SUCCESS;
    if (q != cudaErrorNotReady) return status_t::ERROR;
    ncclResult_t async_err;
    if (ncclCommGetAsyncError(comm, &async_err) != ncclSuccess) return status_t::ERROR;
    if (async_err != ncclSuccess) {
      ncclCommAbort(comm);
      return status_t::ABORT;
    }
  }
}
}  // namespace detail
}  // namespace comms
}  // namespace raft
