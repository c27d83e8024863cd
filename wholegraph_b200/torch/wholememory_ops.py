"""Functional wrappers of gather / scatter on a raw PyWholeMemoryTensor
(mirror of pylibwholegraph/torch/wholememory_ops.py:24-78; WholeMemoryTensor.gather/scatter are the object form)."""
import torch

from .. import binding as wmb
from .utils import wholememory_dtype_to_torch_dtype
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor


def _check_indices(indices_tensor: torch.Tensor):
    assert indices_tensor.dim() == 1
    assert indices_tensor.dtype in (torch.int32, torch.int64)


def wholememory_gather_forward_functor(wholememory_tensor: wmb.PyWholeMemoryTensor, indices_tensor: torch.Tensor,
                                       requires_grad=False, torch_output_dtype=None):
    """out[i, :] = table[indices[i], :] as a new cuda tensor (dtype of the table unless torch_output_dtype is given)."""
    _check_indices(indices_tensor)
    if torch_output_dtype is None:
        torch_output_dtype = wholememory_dtype_to_torch_dtype(wholememory_tensor.dtype)
    output_tensor = torch.empty([indices_tensor.shape[0], wholememory_tensor.shape[1]], device="cuda",
                                dtype=torch_output_dtype, requires_grad=requires_grad)
    wmb.wholememory_gather_op(wholememory_tensor, wrap_torch_tensor(indices_tensor), wrap_torch_tensor(output_tensor),
                              get_wholegraph_env_fns(), get_stream())
    return output_tensor


def wholememory_scatter_functor(input_tensor: torch.Tensor, indices_tensor: torch.Tensor,
                                wholememory_tensor: wmb.PyWholeMemoryTensor):
    """table[indices[i], :] = input[i, :]; returns None."""
    _check_indices(indices_tensor)
    wmb.wholememory_scatter_op(wrap_torch_tensor(input_tensor), wrap_torch_tensor(indices_tensor), wholememory_tensor,
                               get_wholegraph_env_fns(), get_stream())
