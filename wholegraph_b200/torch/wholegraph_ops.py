"""Graph ops on CSR WholeMemory (mirror of pylibwholegraph/torch/wholegraph_ops.py:26-110)."""
import random
from typing import Union

import torch

from .. import binding as wmb
from .wholegraph_env import TorchMemoryContext, get_stream, get_wholegraph_env_fns, wrap_torch_tensor


def unweighted_sample_without_replacement(wm_csr_row_ptr_tensor: wmb.PyWholeMemoryTensor,
                                          wm_csr_col_ptr_tensor: wmb.PyWholeMemoryTensor,
                                          center_nodes_tensor: torch.Tensor,
                                          max_sample_count: int,
                                          random_seed: Union[int, None] = None,
                                          need_center_local_output: bool = False,
                                          need_edge_output: bool = False):
    """Unweighted neighborhood sample in CSR WholeGraph.

    Returns (sample_offset[n+1], dst_nodes[S]) plus center-local ids and/or edge global ids when asked for.
    Variable-size outputs are allocated by the library through the torch env functions."""
    assert wm_csr_row_ptr_tensor.dim() == 1
    assert wm_csr_col_ptr_tensor.dim() == 1
    assert center_nodes_tensor.dim() == 1
    if random_seed is None:
        random_seed = random.getrandbits(64)
    output_sample_offset_tensor = torch.empty(center_nodes_tensor.shape[0] + 1, device="cuda", dtype=torch.int)
    output_dest_context = TorchMemoryContext()
    output_center_localid_context = TorchMemoryContext() if need_center_local_output else None
    output_edge_gid_context = TorchMemoryContext() if need_edge_output else None
    wmb.csr_unweighted_sample_without_replacement(
        wm_csr_row_ptr_tensor,
        wm_csr_col_ptr_tensor,
        wrap_torch_tensor(center_nodes_tensor),
        max_sample_count,
        wrap_torch_tensor(output_sample_offset_tensor),
        output_dest_context.get_c_context(),
        output_center_localid_context.get_c_context() if output_center_localid_context else 0,
        output_edge_gid_context.get_c_context() if output_edge_gid_context else 0,
        random_seed,
        get_wholegraph_env_fns(),
        get_stream(),
    )
    result = [output_sample_offset_tensor, output_dest_context.get_tensor()]
    if need_center_local_output:
        result.append(output_center_localid_context.get_tensor())
    if need_edge_output:
        result.append(output_edge_gid_context.get_tensor())
    for c in (output_dest_context, output_center_localid_context, output_edge_gid_context):
        if c is not None:
            c.free()
    return tuple(result)


def weighted_sample_without_replacement(wm_csr_row_ptr_tensor: wmb.PyWholeMemoryTensor,
                                        wm_csr_col_ptr_tensor: wmb.PyWholeMemoryTensor,
                                        wm_csr_weight_ptr_tensor: wmb.PyWholeMemoryTensor,
                                        center_nodes_tensor: torch.Tensor,
                                        max_sample_count: int,
                                        random_seed: Union[int, None] = None,
                                        need_center_local_output: bool = False,
                                        need_edge_output: bool = False):
    """Weighted neighborhood sample in CSR WholeGraph (probability of keeping an edge grows with its weight; A-Res)."""
    assert wm_csr_row_ptr_tensor.dim() == 1
    assert wm_csr_col_ptr_tensor.dim() == 1
    assert wm_csr_weight_ptr_tensor.dim() == 1
    assert wm_csr_weight_ptr_tensor.shape[0] == wm_csr_col_ptr_tensor.shape[0]
    assert center_nodes_tensor.dim() == 1
    if random_seed is None:
        random_seed = random.getrandbits(64)
    output_sample_offset_tensor = torch.empty(center_nodes_tensor.shape[0] + 1, device="cuda", dtype=torch.int)
    output_dest_context = TorchMemoryContext()
    output_center_localid_context = TorchMemoryContext() if need_center_local_output else None
    output_edge_gid_context = TorchMemoryContext() if need_edge_output else None
    wmb.csr_weighted_sample_without_replacement(
        wm_csr_row_ptr_tensor,
        wm_csr_col_ptr_tensor,
        wm_csr_weight_ptr_tensor,
        wrap_torch_tensor(center_nodes_tensor),
        max_sample_count,
        wrap_torch_tensor(output_sample_offset_tensor),
        output_dest_context.get_c_context(),
        output_center_localid_context.get_c_context() if output_center_localid_context else 0,
        output_edge_gid_context.get_c_context() if output_edge_gid_context else 0,
        random_seed,
        get_wholegraph_env_fns(),
        get_stream(),
    )
    result = [output_sample_offset_tensor, output_dest_context.get_tensor()]
    if need_center_local_output:
        result.append(output_center_localid_context.get_tensor())
    if need_edge_output:
        result.append(output_edge_gid_context.get_tensor())
    for c in (output_dest_context, output_center_localid_context, output_edge_gid_context):
        if c is not None:
            c.free()
    return tuple(result)


def generate_exponential_distribution_negative_float_cpu(random_seed: int, sub_sequence: int, output_random_value_count: int):
    output = torch.empty((output_random_value_count,), dtype=torch.float)
    wmb.host_generate_exponential_distribution_negative_float(random_seed, sub_sequence, wrap_torch_tensor(output))
    return output


def generate_random_positive_int_cpu(random_seed: int, sub_sequence: int, output_random_value_count: int) -> torch.Tensor:
    """Host replay of the sampler's random stream (used by tests to rebuild expected samples)."""
    output = torch.empty((output_random_value_count,), dtype=torch.int)
    wmb.host_generate_random_positive_int(random_seed, sub_sequence, wrap_torch_tensor(output))
    return output
