"""Host-side helpers the reference's Python tests are written with (same names and results as
pylibwholegraph/test_utils/test_comm.py): a random CSR graph generator, the "take every neighbor" host model of the
samplers, a host -> WholeMemory copy, the small-int -> enum maps of the parametrized tests and the seeded random row
partition.  Everything here runs on the CPU; nothing is on the product path."""
import numpy as np
import torch

from .. import binding as wmb
from ..torch.dlpack_utils import torch_import_from_dlpack


def gen_csr_graph(graph_node_count, graph_edge_count, neighbor_node_count=None, csr_row_dtype=torch.int64,
                  csr_col_dtype=torch.int32, weight_dtype=torch.float32):
    """A random bipartite graph with exactly graph_edge_count distinct (row, col) edges and weights in [1, 2):
    (csr_row_ptr[nodes + 1], csr_col_ptr[edges] ascending inside each row, csr_weight_ptr[edges])."""
    if neighbor_node_count is None:
        neighbor_node_count = graph_node_count
    cells = graph_node_count * neighbor_node_count
    assert cells >= graph_edge_count
    dense = torch.rand(cells, dtype=weight_dtype, device="cpu") + 1           # every cell a candidate edge with weight in [1, 2)
    dense[torch.randperm(cells, device="cpu")[:cells - graph_edge_count]] = 0  # keep exactly graph_edge_count of them
    dense = dense.reshape(graph_node_count, neighbor_node_count)
    rows, cols = torch.nonzero(dense, as_tuple=True)                           # row-major: grouped by row, columns ascending
    csr_row_ptr = torch.zeros(graph_node_count + 1, dtype=torch.int64)
    csr_row_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=graph_node_count), dim=0)
    assert int(csr_row_ptr[-1]) == graph_edge_count
    return csr_row_ptr.to(csr_row_dtype), cols.to(csr_col_dtype), dense[rows, cols]


def host_get_sample_offset_tensor(host_csr_row_ptr, center_nodes, max_sample_count):
    """int32 [n + 1]: exclusive prefix sums of min(degree, max_sample_count) (all neighbors when max_sample_count <= 0)."""
    row_ptr = host_csr_row_ptr.to(torch.int64)
    centers = center_nodes.to(torch.int64)
    counts = row_ptr[centers + 1] - row_ptr[centers]
    if max_sample_count > 0:
        counts = torch.clamp(counts, max=max_sample_count)
    offsets = torch.zeros(centers.shape[0] + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(counts, dim=0).to(torch.int32)
    return offsets


def host_sample_all_neighbors(host_csr_row_ptr, host_csr_col_ptr, center_nodes, output_sample_offset_tensor, col_id_dtype,
                              total_sample_count):
    """What the samplers return when every neighbor is taken: (offsets, neighbor ids, center-local ids int32, edge ids int64)."""
    row_ptr = host_csr_row_ptr.to(torch.int64)
    centers = center_nodes.to(torch.int64)
    degrees = row_ptr[centers + 1] - row_ptr[centers]
    local_ids = torch.repeat_interleave(torch.arange(centers.shape[0], dtype=torch.int64), degrees)
    first_out = output_sample_offset_tensor.to(torch.int64)[:-1]
    edge_ids = row_ptr[centers][local_ids] + (torch.arange(int(degrees.sum()), dtype=torch.int64) - first_out[local_ids])
    assert edge_ids.shape[0] == total_sample_count
    return (output_sample_offset_tensor, host_csr_col_ptr[edge_ids].to(col_id_dtype), local_ids.to(torch.int32), edge_ids)


def copy_host_1D_tensor_to_wholememory(wm_array, host_tensor, world_rank, world_size, wm_comm):
    """Every rank copies its slice of the 1-D host tensor into its partition of wm_array, then all ranks barrier."""
    local, first = wm_array.get_local_tensor(torch_import_from_dlpack, wmb.WholeMemoryMemoryLocation.MlDevice, world_rank)
    count = wm_array.get_local_entry_count()
    assert local.dim() == 1 and local.shape[0] == count and first == wm_array.get_local_entry_start()
    local.copy_(host_tensor[first:first + count])
    wm_comm.barrier()


_DATATYPES = (wmb.WholeMemoryDataType.DtInt, wmb.WholeMemoryDataType.DtInt64, wmb.WholeMemoryDataType.DtFloat,
              wmb.WholeMemoryDataType.DtDouble)
_LOCATIONS = (wmb.WholeMemoryMemoryLocation.MlHost, wmb.WholeMemoryMemoryLocation.MlDevice)
_MEMORY_TYPES = (wmb.WholeMemoryMemoryType.MtContinuous, wmb.WholeMemoryMemoryType.MtChunked,
                 wmb.WholeMemoryMemoryType.MtDistributed, wmb.WholeMemoryMemoryType.MtHierarchy)


def _pick(table, value, what):
    if not 0 <= value < len(table):
        raise ValueError("invalid %s value" % what)
    return table[value]


def int_to_wholememory_datatype(value: int):
    """0 int32, 1 int64, 2 float, 3 double (the parametrize ids of the reference tests)"""
    return _pick(_DATATYPES, value, "int_to_wholememory_datatype")


def int_to_wholememory_location(value: int):
    """0 host, 1 device"""
    return _pick(_LOCATIONS, value, "int_to_wholememory_location")


def int_to_wholememory_type(value: int):
    """0 continuous, 1 chunked, 2 distributed, 3 hierarchy"""
    return _pick(_MEMORY_TYPES, value, "int_to_wholememory_type")


def random_partition(total_entry_count: int, world_size: int) -> np.array:
    """Uneven but balanced row counts per rank (shares drawn from U(90, 100) with the fixed seed 42, so every rank computes
    the same partition); the rounding remainder goes to rank 0."""
    np.random.seed(42)
    shares = np.random.uniform(90, 100, size=world_size)
    partition = (shares / shares.sum() * total_entry_count).astype(np.uintp)
    partition[0] += total_entry_count - partition.sum()
    return partition
