"""GraphStructure: one relation of a graph as CSR arrays in WholeMemory, plus node / edge attribute tensors and the
one-hop and multi-hop neighbor samplers built on them.  Public names follow pylibwholegraph/torch/graph_structure.py
(the GNN-model glue around it is out of scope)."""
from typing import List, Union

import torch

from . import graph_ops, wholegraph_ops
from .tensor import WholeMemoryTensor


class GraphStructure(object):
    def __init__(self):
        super().__init__()
        self.csr_row_ptr = None   # WholeMemoryTensor int64 [nodes + 1]
        self.csr_col_ind = None   # WholeMemoryTensor int32 | int64 [edges]
        self.node_count = 0
        self.edge_count = 0
        self.node_attributes = {}
        self.edge_attributes = {}

    def set_csr_graph(self, csr_row_ptr: WholeMemoryTensor, csr_col_ind: WholeMemoryTensor):
        """row_ptr: int64 [nodes+1]; col_ind: int32|int64 [edges]."""
        assert csr_row_ptr.dim() == 1 and csr_col_ind.dim() == 1
        assert csr_row_ptr.dtype == torch.int64
        assert csr_col_ind.dtype in (torch.int32, torch.int64)
        assert csr_row_ptr.shape[0] > 1
        self.csr_row_ptr, self.csr_col_ind = csr_row_ptr, csr_col_ind
        self.node_count, self.edge_count = csr_row_ptr.shape[0] - 1, csr_col_ind.shape[0]

    @staticmethod
    def _register(table: dict, expected_rows: int, attr_name: str, attr_tensor: WholeMemoryTensor):
        assert attr_name not in table
        assert attr_tensor.shape[0] == expected_rows
        table[attr_name] = attr_tensor

    def set_node_attribute(self, attr_name: str, attr_tensor: WholeMemoryTensor):
        """One row per node."""
        self._register(self.node_attributes, self.node_count, attr_name, attr_tensor)

    def set_edge_attribute(self, attr_name: str, attr_tensor: WholeMemoryTensor):
        """One row per edge, in CSR order (e.g. the weights used by weighted sampling)."""
        self._register(self.edge_attributes, self.edge_count, attr_name, attr_tensor)

    def unweighted_sample_without_replacement_one_hop(self, center_nodes_tensor: torch.Tensor, max_sample_count: int, *,
                                                      random_seed: Union[int, None] = None,
                                                      need_center_local_output: bool = False, need_edge_output: bool = False):
        """(offsets[n+1], neighbor ids[S]) + center-local ids and/or edge ids when asked for."""
        return wholegraph_ops.unweighted_sample_without_replacement(self.csr_row_ptr.wmb_tensor, self.csr_col_ind.wmb_tensor,
                                                                    center_nodes_tensor, max_sample_count, random_seed,
                                                                    need_center_local_output, need_edge_output)

    def weighted_sample_without_replacement_one_hop(self, weight_name: str, center_nodes_tensor: torch.Tensor, max_sample_count: int, *,
                                                    random_seed: Union[int, None] = None,
                                                    need_center_local_output: bool = False, need_edge_output: bool = False):
        """Same outputs, neighbors kept with probability growing with the edge attribute `weight_name`."""
        assert weight_name in self.edge_attributes
        weights = self.edge_attributes[weight_name]
        return wholegraph_ops.weighted_sample_without_replacement(self.csr_row_ptr.wmb_tensor, self.csr_col_ind.wmb_tensor,
                                                                  weights.wmb_tensor, center_nodes_tensor, max_sample_count,
                                                                  random_seed, need_center_local_output, need_edge_output)

    def multilayer_sample_without_replacement(self, node_ids: torch.Tensor, max_neighbors: List[int],
                                              weight_name: Union[str, None] = None, random_seed: Union[int, None] = None):
        """Sample len(max_neighbors) hops outwards from the seed nodes `node_ids`.

        Layer numbering follows the reference (graph_structure.py:160-196): the seeds are target_gids[hops] and each hop
        fills the next LOWER layer, hop number h (counting from the seeds) using fan-out max_neighbors[h] and seed
        random_seed + (hops - 1 - h).  After each hop the frontier is targets ++ new neighbors (append_unique), so layer
        i's targets are a prefix of layer i - 1's.  Returns (target_gids[hops+1], edge_indice[hops], csr_row_ptr[hops],
        csr_col_ind[hops]); edge_indice[i] is a [2, E_i] tensor (neighbor index in layer i's frontier, center index)."""
        hops = len(max_neighbors)
        target_gids = [None] * hops + [node_ids]
        edge_indice, csr_row_ptr, csr_col_ind = [None] * hops, [None] * hops, [None] * hops
        for layer in reversed(range(hops)):
            centers = target_gids[layer + 1]
            fanout = max_neighbors[hops - 1 - layer]
            seed = None if random_seed is None else random_seed + layer
            if weight_name is None:
                offsets, neighbors, center_lids = self.unweighted_sample_without_replacement_one_hop(
                    centers, fanout, random_seed=seed, need_center_local_output=True)
            else:
                offsets, neighbors, center_lids = self.weighted_sample_without_replacement_one_hop(
                    weight_name, centers, fanout, random_seed=seed, need_center_local_output=True)
            frontier, neighbor_pos = graph_ops.append_unique(centers, neighbors, need_neighbor_raw_to_unique=True)
            target_gids[layer] = frontier
            csr_row_ptr[layer], csr_col_ind[layer] = offsets, neighbor_pos
            edge_indice[layer] = torch.stack([neighbor_pos.reshape(-1), center_lids.reshape(-1)])
        return target_gids, edge_indice, csr_row_ptr, csr_col_ind
