"""Graph helper ops between sampling hops (mirror of pylibwholegraph/torch/graph_ops.py)."""
import torch

from .. import binding as wmb
from .wholegraph_env import TorchMemoryContext, get_stream, get_wholegraph_env_fns, wrap_torch_tensor


def append_unique(target_node_tensor: torch.Tensor, neighbor_node_tensor: torch.Tensor, need_neighbor_raw_to_unique: bool = False):
    """unique = targets ++ (distinct neighbors not among targets, in order of first occurrence);
    e.g. targets [3, 11, 2, 10], neighbors [4, 5, 2, 11, 6, 9, 10, 5] -> unique [3, 11, 2, 10, 4, 5, 6, 9],
    neighbor_raw_to_unique_mapping [4, 5, 2, 1, 6, 7, 3, 5] (the reference leaves the order of the appended part unspecified)."""
    assert target_node_tensor.dim() == 1
    assert neighbor_node_tensor.dim() == 1
    assert target_node_tensor.is_cuda
    assert neighbor_node_tensor.is_cuda
    output_unique_node_context = TorchMemoryContext()
    mapping = None
    if need_neighbor_raw_to_unique:
        mapping = torch.empty(neighbor_node_tensor.shape[0], device="cuda", dtype=torch.int)
    wmb.append_unique(wrap_torch_tensor(target_node_tensor), wrap_torch_tensor(neighbor_node_tensor),
                      output_unique_node_context.get_c_context(), wrap_torch_tensor(mapping), get_wholegraph_env_fns(), get_stream())
    unique = output_unique_node_context.get_tensor()
    output_unique_node_context.free()
    return (unique, mapping) if need_neighbor_raw_to_unique else unique


def add_csr_self_loop(csr_row_ptr_tensor: torch.Tensor, csr_col_ptr_tensor: torch.Tensor):
    """Add one self edge (placed first) to every row of a sampled int32 CSR graph; existing self loops are not detected."""
    assert csr_row_ptr_tensor.dim() == 1
    assert csr_col_ptr_tensor.dim() == 1
    assert csr_row_ptr_tensor.is_cuda
    assert csr_col_ptr_tensor.is_cuda
    if csr_row_ptr_tensor.dtype != torch.int32:
        csr_row_ptr_tensor = csr_row_ptr_tensor.int()
    if csr_col_ptr_tensor.dtype != torch.int32:
        csr_col_ptr_tensor = csr_col_ptr_tensor.int()
    rows = csr_row_ptr_tensor.shape[0] - 1
    out_row = torch.empty(csr_row_ptr_tensor.shape[0], device="cuda", dtype=torch.int)
    out_col = torch.empty(csr_col_ptr_tensor.shape[0] + rows, device="cuda", dtype=torch.int)
    wmb.add_csr_self_loop(wrap_torch_tensor(csr_row_ptr_tensor), wrap_torch_tensor(csr_col_ptr_tensor),
                          wrap_torch_tensor(out_row), wrap_torch_tensor(out_col), get_stream())
    return out_row, out_col
