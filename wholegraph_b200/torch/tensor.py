"""WholeMemoryTensor: a 1-D / 2-D strided tensor row-sharded over a communicator, with the public methods of
pylibwholegraph/torch/tensor.py (gather / scatter / sub-tensor / local, global and per-rank views / part-file I/O)."""
from typing import List, Union

import torch

from .. import binding as wmb
from .comm import WholeMemoryCommunicator
from .utils import (count_file_entries, get_part_file_list, get_part_file_name, str_to_wmb_wholememory_location,
                    str_to_wmb_wholememory_memory_type, torch_dtype_to_wholememory_dtype, wholememory_dtype_to_torch_dtype)
from .wholegraph_env import current_output_device, get_stream, get_wholegraph_env_fns, wrap_torch_tensor

WholeMemoryMemoryType = wmb.WholeMemoryMemoryType
WholeMemoryMemoryLocation = wmb.WholeMemoryMemoryLocation


def _view_args(host_view: bool):
    """(location, device id) the binding's view getters expect: the host, or the current CUDA device."""
    if host_view:
        return WholeMemoryMemoryLocation.MlHost, -1
    return WholeMemoryMemoryLocation.MlDevice, torch.cuda.current_device()


def _as_list(filelist: Union[List[str], str]) -> List[str]:
    return [filelist] if isinstance(filelist, str) else filelist


class WholeMemoryTensor(object):
    """Python handle of one wholememory_tensor_t."""

    def __init__(self, wmb_tensor: wmb.PyWholeMemoryTensor):
        self.wmb_tensor = wmb_tensor

    # ---- description
    @property
    def dtype(self):
        return wholememory_dtype_to_torch_dtype(self.wmb_tensor.dtype)

    @property
    def shape(self):
        return self.wmb_tensor.shape

    def dim(self):
        return self.wmb_tensor.dim()

    def stride(self):
        return self.wmb_tensor.stride()

    def storage_offset(self):
        return self.wmb_tensor.storage_offset()

    def get_comm(self):
        return WholeMemoryCommunicator(self.wmb_tensor.get_wholememory_handle().get_communicator())

    # ---- the hot path
    def gather(self, indice: torch.Tensor, *, force_dtype: Union[torch.dtype, None] = None):
        """rows[i, :] = self[indice[i], :] as a new tensor on the current CUDA device (dtype of the table unless forced)."""
        assert indice.dim() == 1
        rows = torch.empty([indice.shape[0], self.shape[1]], device=current_output_device(),
                           dtype=self.dtype if force_dtype is None else force_dtype, requires_grad=False)
        wmb.wholememory_gather_op(self.wmb_tensor, wrap_torch_tensor(indice), wrap_torch_tensor(rows),
                                  get_wholegraph_env_fns(), get_stream())
        return rows

    def scatter(self, input_tensor: torch.Tensor, indice: torch.Tensor):
        """self[indice[i], :] = input_tensor[i, :]"""
        assert indice.dim() == 1 and input_tensor.dim() == 2
        assert input_tensor.shape[0] == indice.shape[0]
        assert input_tensor.shape[1] == self.shape[1]
        wmb.wholememory_scatter_op(wrap_torch_tensor(input_tensor), wrap_torch_tensor(indice), self.wmb_tensor,
                                   get_wholegraph_env_fns(), get_stream())

    # ---- views
    def get_sub_tensor(self, starts, ends):
        """[starts, ends) per dim; -1 in ends means "to the last element"."""
        return WholeMemoryTensor(self.wmb_tensor.get_sub_tensor(starts, ends))

    def get_local_tensor(self, host_view: bool = False):
        """(torch view of this rank's rows, index of its first row)"""
        return self.wmb_tensor.get_local_tensor(*_view_args(host_view))

    def get_global_tensor(self, host_view: bool = False):
        """torch view of the whole tensor (CONTINUOUS memory, or CHUNKED host memory)"""
        return self.wmb_tensor.get_global_tensor(*_view_args(host_view))

    def get_all_chunked_tensor(self, host_view: bool = False):
        """(one torch view per rank, first row of each)"""
        return self.wmb_tensor.get_all_chunked_tensor(*_view_args(host_view))

    # ---- raw binary part files: "<prefix>_part_<i>_of_<n>"
    def from_filelist(self, filelist: Union[List[str], str], round_robin_size: int = 0):
        self.wmb_tensor.from_filelist(_as_list(filelist), round_robin_size)

    def from_file_prefix(self, file_prefix: str, part_count: Union[int, None] = None):
        """Load the part files of `file_prefix`; part_count defaults to the communicator size."""
        parts = self.get_comm().get_size() if part_count is None else part_count
        self.from_filelist(get_part_file_list(file_prefix, parts))

    def local_to_file(self, filename: str):
        """Store this rank's rows; collective, every rank passes its own file name."""
        self.wmb_tensor.to_file(filename)

    def to_file_prefix(self, file_prefix: str):
        comm = self.get_comm()
        self.local_to_file(get_part_file_name(file_prefix, comm.get_rank(), comm.get_size()))


def create_wholememory_tensor(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str, sizes: List[int],
                              dtype: torch.dtype, strides: List[int],
                              tensor_entry_partition: Union[List[int], None] = None):
    """Create an empty WholeMemory tensor of 1 or 2 dims.  strides=None means packed rows;
    tensor_entry_partition[i] = rows owned by rank i (default: equal split)."""
    if len(sizes) not in (1, 2):
        raise ValueError("Only dim 1 or 2 is supported now.")
    if strides is None:
        strides = [1] if len(sizes) == 1 else [sizes[1], 1]
    else:
        assert len(strides) == len(sizes) and strides[-1] == 1
        assert len(sizes) == 1 or strides[0] >= sizes[1]
    desc = wmb.PyWholeMemoryTensorDescription()
    desc.set_shape(sizes)
    desc.set_stride(strides)
    desc.set_dtype(torch_dtype_to_wholememory_dtype(dtype))
    return WholeMemoryTensor(wmb.create_wholememory_tensor(desc, comm.wmb_comm, str_to_wmb_wholememory_memory_type(memory_type),
                                                           str_to_wmb_wholememory_location(memory_location), tensor_entry_partition))


def create_wholememory_tensor_from_filelist(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str,
                                            filelist: Union[List[str], str], dtype: torch.dtype, last_dim_size: int = 0,
                                            last_dim_strides: int = -1, tensor_entry_partition: Union[List[int], None] = None):
    """Create a WholeMemory tensor sized from, and filled with, raw binary files: last_dim_size 0 gives a 1-D array of
    `dtype`, last_dim_size > 0 a matrix with that many columns (row stride last_dim_strides, packed by default)."""
    filelist = _as_list(filelist)
    element_bytes = torch.tensor([], dtype=dtype).element_size()
    entries = count_file_entries(filelist, element_bytes * max(last_dim_size, 1))
    if last_dim_size == 0:
        sizes, strides = [entries], [1]
    else:
        sizes, strides = [entries, last_dim_size], [last_dim_size if last_dim_strides == -1 else last_dim_strides, 1]
    wm_tensor = create_wholememory_tensor(comm, memory_type, memory_location, sizes, dtype, strides, tensor_entry_partition)
    wm_tensor.from_filelist(filelist)
    return wm_tensor


def destroy_wholememory_tensor(wm_tensor: WholeMemoryTensor):
    wmb.destroy_wholememory_tensor(wm_tensor.wmb_tensor)
    wm_tensor.wmb_tensor = None
