"""WholeMemoryTensor (mirror of pylibwholegraph/torch/tensor.py:30-330): same methods, same semantics."""
from typing import List, Union

import torch

from .. import binding as wmb
from .comm import WholeMemoryCommunicator
from .utils import (get_file_size, get_part_file_list, get_part_file_name, str_to_wmb_wholememory_location,
                    str_to_wmb_wholememory_memory_type, torch_dtype_to_wholememory_dtype, wholememory_dtype_to_torch_dtype)
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

WholeMemoryMemoryType = wmb.WholeMemoryMemoryType
WholeMemoryMemoryLocation = wmb.WholeMemoryMemoryLocation


class WholeMemoryTensor(object):
    r"""WholeMemory Tensor"""

    def __init__(self, wmb_tensor: wmb.PyWholeMemoryTensor):
        self.wmb_tensor = wmb_tensor

    @property
    def dtype(self):
        return wholememory_dtype_to_torch_dtype(self.wmb_tensor.dtype)

    def dim(self):
        return self.wmb_tensor.dim()

    @property
    def shape(self):
        return self.wmb_tensor.shape

    def stride(self):
        return self.wmb_tensor.stride()

    def storage_offset(self):
        return self.wmb_tensor.storage_offset()

    def get_comm(self):
        return WholeMemoryCommunicator(self.wmb_tensor.get_wholememory_handle().get_communicator())

    def gather(self, indice: torch.Tensor, *, force_dtype: Union[torch.dtype, None] = None):
        assert indice.dim() == 1
        embedding_dim = self.shape[1]
        embedding_count = indice.shape[0]
        current_cuda_device = "cuda:%d" % (torch.cuda.current_device(),)
        output_dtype = force_dtype if force_dtype is not None else self.dtype
        output_tensor = torch.empty([embedding_count, embedding_dim], device=current_cuda_device, dtype=output_dtype,
                                    requires_grad=False)
        wmb.wholememory_gather_op(self.wmb_tensor, wrap_torch_tensor(indice), wrap_torch_tensor(output_tensor),
                                  get_wholegraph_env_fns(), get_stream())
        return output_tensor

    def scatter(self, input_tensor: torch.Tensor, indice: torch.Tensor):
        assert indice.dim() == 1
        assert input_tensor.dim() == 2
        assert indice.shape[0] == input_tensor.shape[0]
        assert input_tensor.shape[1] == self.shape[1]
        wmb.wholememory_scatter_op(wrap_torch_tensor(input_tensor), wrap_torch_tensor(indice), self.wmb_tensor,
                                   get_wholegraph_env_fns(), get_stream())

    def get_sub_tensor(self, starts, ends):
        """[starts, ends) per dim; -1 in ends means "to the last element"."""
        return WholeMemoryTensor(self.wmb_tensor.get_sub_tensor(starts, ends))

    def get_local_tensor(self, host_view: bool = False):
        """(torch view of this rank's rows, first-row offset)"""
        if host_view:
            return self.wmb_tensor.get_local_tensor(WholeMemoryMemoryLocation.MlHost, -1)
        return self.wmb_tensor.get_local_tensor(WholeMemoryMemoryLocation.MlDevice, torch.cuda.current_device())

    def get_global_tensor(self, host_view: bool = False):
        if host_view:
            return self.wmb_tensor.get_global_tensor(WholeMemoryMemoryLocation.MlHost, -1)
        return self.wmb_tensor.get_global_tensor(WholeMemoryMemoryLocation.MlDevice, torch.cuda.current_device())

    def get_all_chunked_tensor(self, host_view: bool = False):
        if host_view:
            return self.wmb_tensor.get_all_chunked_tensor(WholeMemoryMemoryLocation.MlHost, -1)
        return self.wmb_tensor.get_all_chunked_tensor(WholeMemoryMemoryLocation.MlDevice, torch.cuda.current_device())

    def from_filelist(self, filelist: Union[List[str], str], round_robin_size: int = 0):
        if isinstance(filelist, str):
            filelist = [filelist]
        self.wmb_tensor.from_filelist(filelist, round_robin_size)

    def from_file_prefix(self, file_prefix: str, part_count: Union[int, None] = None):
        """Load from files named "%s_part_%d_of_%d" % (prefix, part_id, part_count)."""
        if part_count is None:
            part_count = self.get_comm().get_size()
        self.from_filelist(get_part_file_list(file_prefix, part_count))

    def local_to_file(self, filename: str):
        """Store this rank's rows; all ranks call it together with different file names."""
        self.wmb_tensor.to_file(filename)

    def to_file_prefix(self, file_prefix: str):
        wm_comm = self.get_comm()
        self.local_to_file(get_part_file_name(file_prefix, wm_comm.get_rank(), wm_comm.get_size()))


def create_wholememory_tensor(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str, sizes: List[int],
                              dtype: torch.dtype, strides: List[int],
                              tensor_entry_partition: Union[List[int], None] = None):
    """Create an empty WholeMemory tensor (dim 1 or 2).  tensor_entry_partition[i] = rows owned by rank i."""
    dim = len(sizes)
    if dim < 1 or dim > 2:
        raise ValueError("Only dim 1 or 2 is supported now.")
    if strides is None:
        strides = [1] * dim
        strides[0] = sizes[1] if dim == 2 else 1
    else:
        assert len(strides) == dim
        assert strides[-1] == 1
        if dim == 2:
            assert strides[0] >= sizes[1]
    td = wmb.PyWholeMemoryTensorDescription()
    td.set_shape(sizes)
    td.set_stride(strides)
    td.set_dtype(torch_dtype_to_wholememory_dtype(dtype))
    wm_memory_type = str_to_wmb_wholememory_memory_type(memory_type)
    wm_location = str_to_wmb_wholememory_location(memory_location)
    return WholeMemoryTensor(
        wmb.create_wholememory_tensor(td, comm.wmb_comm, wm_memory_type, wm_location, tensor_entry_partition))


def create_wholememory_tensor_from_filelist(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str,
                                            filelist: Union[List[str], str], dtype: torch.dtype, last_dim_size: int = 0,
                                            last_dim_strides: int = -1, tensor_entry_partition: Union[List[int], None] = None):
    """Create a WholeMemory tensor sized from, and filled with, a list of raw binary files
    (last_dim_size 0 -> 1-D array, > 0 -> matrix with that many columns)."""
    if isinstance(filelist, str):
        filelist = [filelist]
    element_size = torch.tensor([], dtype=dtype).element_size()
    if last_dim_strides == -1:
        last_dim_strides = last_dim_size if last_dim_size > 0 else 1
    file_entry_size = element_size * last_dim_size if last_dim_size > 0 else element_size
    total_file_size = 0
    for filename in filelist:
        file_size = get_file_size(filename)
        if file_size % file_entry_size != 0:
            raise ValueError("File %s size is %d not mutlple of %d" % (filename, file_size, file_entry_size))
        total_file_size += file_size
    total_entry_count = total_file_size // file_entry_size
    if last_dim_size == 0:
        sizes, strides = [total_entry_count], [1]
    else:
        sizes, strides = [total_entry_count, last_dim_size], [last_dim_strides, 1]
    wm_tensor = create_wholememory_tensor(comm, memory_type, memory_location, sizes, dtype, strides, tensor_entry_partition)
    wm_tensor.from_filelist(filelist)
    return wm_tensor


def destroy_wholememory_tensor(wm_tensor: WholeMemoryTensor):
    wmb.destroy_wholememory_tensor(wm_tensor.wmb_tensor)
    wm_tensor.wmb_tensor = None
