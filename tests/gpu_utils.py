"""Helpers shared by the -m gpu tests (single process, world_size 1)."""
import numpy as np
import torch

import wholegraph_b200.binding as wmb
from oracle import oracle as O
from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

WM_OF = {O.DT_FLOAT: wmb.DtFloat, O.DT_HALF: wmb.DtHalf, O.DT_DOUBLE: wmb.DtDouble, O.DT_BF16: wmb.DtBF16,
         O.DT_INT: wmb.DtInt, O.DT_INT64: wmb.DtInt64, O.DT_INT16: wmb.DtInt16, O.DT_INT8: wmb.DtInt8}
TORCH_OF = {O.DT_FLOAT: torch.float32, O.DT_HALF: torch.float16, O.DT_DOUBLE: torch.float64, O.DT_BF16: torch.bfloat16,
            O.DT_INT: torch.int32, O.DT_INT64: torch.int64, O.DT_INT16: torch.int16, O.DT_INT8: torch.int8}
MT = {"continuous": wmb.MtContinuous, "chunked": wmb.MtChunked, "distributed": wmb.MtDistributed}
ML = {"cuda": wmb.MlDevice, "cpu": wmb.MlHost}

_comm = None


def single_comm():
    """The world_size-1 communicator the single-process GPU tests share.  Re-created when something finalized the
    library in between (wholememory_finalize destroys every communicator; a stale one is refused with INVALID_INPUT)."""
    global _comm
    if _comm is not None:
        try:
            _comm.get_rank()
        except Exception:
            _comm = None
    if _comm is None:
        wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
        torch.cuda.set_device(0)
        _comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    return _comm


def np_to_torch(a, dt):
    """numpy array (bf16 as uint16 bits) -> torch tensor of the real dtype, on the host."""
    if a.size == 0:  # from_numpy of an empty array reports zero strides; build it natively instead
        return torch.empty(a.shape, dtype=TORCH_OF[dt])
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.view(torch.bfloat16) if dt == O.DT_BF16 else t


def torch_to_np(t, dt):
    t = t.detach().cpu().contiguous().clone()  # clone: host WholeMemory views alias memory that is unmapped on destroy
    return t.view(torch.int16).numpy().view(np.uint16) if dt == O.DT_BF16 else t.numpy()


def idx_to_cuda(idx_np):
    """numpy index array -> cuda tensor; an EMPTY from_numpy tensor reports stride 0, which the C ABI (like the
    reference's) rejects for 1-D arrays, so empty batches are created natively."""
    if idx_np.shape[0] == 0:
        return torch.empty(0, dtype=torch.int64 if idx_np.dtype == np.int64 else torch.int32, device="cuda")
    return torch.from_numpy(idx_np).cuda()


def random_table(rng, dt, rows, stride):
    if dt in O.INT_DTS:
        info = np.iinfo(O.NP_OF[dt])
        return rng.integers(info.min, info.max, size=(rows, stride), dtype=O.NP_OF[dt], endpoint=True)
    if dt == O.DT_BF16:
        return O._f32_to_bf16(rng.standard_normal((rows, stride)).astype(np.float32) * 50)
    scale = rng.choice([1e-5, 1.0, 300.0, 6e4], size=(rows, stride))
    with np.errstate(over="ignore"):
        return (rng.standard_normal((rows, stride)) * scale).astype(O.NP_OF[dt])


def create_table(comm, mem_type, location, dt, rows, cols, stride, partition=None):
    """(PyWholeMemoryTensor, device/host torch view of ALL its elements as [rows, stride]) at world_size 1."""
    t = wmb.create_wholememory_matrix(WM_OF[dt], rows, cols, stride, comm, MT[mem_type], ML[location], partition)
    h = t.get_wholememory_handle()
    view_loc = wmb.MlDevice if location == "cuda" else wmb.MlHost
    flat, off = h.get_local_flatten_tensor(WM_OF[dt], view_loc, torch.cuda.current_device())
    assert off == 0
    return t, flat.reshape(rows, stride)


def gather(table, idx_t, out_t, sms=-1):
    wmb.wholememory_gather_op(table, wrap_torch_tensor(idx_t), wrap_torch_tensor(out_t), get_wholegraph_env_fns(),
                              get_stream(), sms)


def scatter(inp_t, idx_t, table, sms=-1):
    wmb.wholememory_scatter_op(wrap_torch_tensor(inp_t), wrap_torch_tensor(idx_t), table, get_wholegraph_env_fns(),
                               get_stream(), sms)
