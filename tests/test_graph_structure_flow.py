"""GraphStructure.multilayer_sample_without_replacement: layer bookkeeping checked on CPU with the two GPU ops it chains
(one-hop sampling, append_unique) replaced by small host implementations.  Contract (reference graph_structure.py:140-196):
seeds are the LAST target layer; hop h from the seeds uses fan-out max_neighbors[h] and seed random_seed + (hops-1-h); every
layer's targets are a prefix of the next lower layer's; csr_row_ptr / csr_col_ind / edge_indice of a layer describe the
sampled block between that layer's centers (rows) and its frontier (columns)."""
import os

import pytest
import torch

import wholegraph_b200.torch as wgth
from wholegraph_b200.torch import graph_structure as gs_mod

ROW_PTR = torch.tensor([0, 3, 5, 9, 9, 12, 14, 18, 20, 21, 24])          # 10 nodes
COL = torch.tensor([1, 2, 3, 0, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 2, 4, 6])


def _fake_one_hop(calls):
    def sample(row_ptr_t, col_t, centers, k, random_seed, need_center_local_output, need_edge_output):
        calls.append((tuple(centers.tolist()), k, random_seed))
        offsets, nbrs, lids = [0], [], []
        for i, c in enumerate(centers.tolist()):
            picked = COL[ROW_PTR[c]:ROW_PTR[c + 1]][:k].tolist()  # deterministic: the first k neighbours
            nbrs += picked
            lids += [i] * len(picked)
            offsets.append(len(nbrs))
        return (torch.tensor(offsets, dtype=torch.int32), torch.tensor(nbrs, dtype=torch.int64), torch.tensor(lids, dtype=torch.int32))
    return sample


def _fake_append_unique(targets, neighbors, need_neighbor_raw_to_unique=False):
    seen = {int(v): i for i, v in enumerate(targets.tolist())}
    uniq = targets.tolist()
    mapping = []
    for v in neighbors.tolist():
        if v not in seen:
            seen[v] = len(uniq)
            uniq.append(v)
        mapping.append(seen[v])
    return torch.tensor(uniq, dtype=torch.int64), torch.tensor(mapping, dtype=torch.int32)


def test_multilayer_sampling_bookkeeping(monkeypatch):
    calls = []
    monkeypatch.setattr(gs_mod.wholegraph_ops, "unweighted_sample_without_replacement", _fake_one_hop(calls))
    monkeypatch.setattr(gs_mod.graph_ops, "append_unique", _fake_append_unique)
    g = wgth.GraphStructure()
    g.csr_row_ptr = g.csr_col_ind = type("T", (), {"wmb_tensor": None})()  # set_csr_graph needs WholeMemory tensors; not needed by the fakes
    seeds = torch.tensor([4, 0, 7])
    fanout = [3, 2]
    target_gids, edge_indice, csr_row_ptr, csr_col_ind = g.multilayer_sample_without_replacement(seeds, fanout, random_seed=100)
    hops = len(fanout)
    assert len(target_gids) == hops + 1 and len(edge_indice) == len(csr_row_ptr) == len(csr_col_ind) == hops
    assert torch.equal(target_gids[hops], seeds)
    # hop 0 (from the seeds) fills layer hops-1 with fan-out max_neighbors[0] and seed random_seed + hops-1; hop 1 the next lower layer
    assert calls[0] == ((4, 0, 7), 3, 101) and calls[1][1:] == (2, 100)
    assert calls[1][0] == tuple(target_gids[1].tolist())
    for layer in range(hops):
        centers, frontier = target_gids[layer + 1], target_gids[layer]
        assert torch.equal(frontier[: centers.shape[0]], centers)             # targets stay in front
        assert frontier.unique().shape[0] == frontier.shape[0]
        offs, cols, ei = csr_row_ptr[layer], csr_col_ind[layer], edge_indice[layer]
        assert offs.shape[0] == centers.shape[0] + 1 and int(offs[-1]) == cols.shape[0]
        assert tuple(ei.shape) == (2, cols.shape[0]) and torch.equal(ei[0], cols)
        k = fanout[hops - 1 - layer]
        for i, c in enumerate(centers.tolist()):
            want = COL[ROW_PTR[c]:ROW_PTR[c + 1]][:k]
            got = frontier[cols[int(offs[i]):int(offs[i + 1])].long()]
            assert torch.equal(got, want)                                      # block row i lists center i's sampled neighbours
            assert torch.equal(ei[1][int(offs[i]):int(offs[i + 1])], torch.full((want.shape[0],), i, dtype=torch.int32))
    # unseeded: every hop passes None on
    calls.clear()
    g.multilayer_sample_without_replacement(seeds, fanout)
    assert [c[2] for c in calls] == [None, None]


REF_GS = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/graph_structure.py"


@pytest.mark.skipif(not os.path.exists(REF_GS), reason="reference tree not present")
@pytest.mark.parametrize("fanout", [[3, 2], [2], [1, 4, 2]])
def test_multilayer_bookkeeping_equals_the_reference_class(monkeypatch, fanout):
    """The reference's GraphStructure (graph_structure.py loaded unchanged; its `from . import graph_ops, wholegraph_ops`
    resolve to this repo's modules through compat/) and this repo's, on the same stand-in ops: identical layer lists."""
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF_GS, "_reference_graph_structure", package="pylibwholegraph.torch")
    calls = []
    monkeypatch.setattr(gs_mod.wholegraph_ops, "unweighted_sample_without_replacement", _fake_one_hop(calls))
    monkeypatch.setattr(gs_mod.graph_ops, "append_unique", _fake_append_unique)
    assert ref_mod.wholegraph_ops is gs_mod.wholegraph_ops and ref_mod.graph_ops is gs_mod.graph_ops  # one set of (patched) ops
    holder = type("T", (), {"wmb_tensor": None})()
    seeds = torch.tensor([4, 0, 7, 9])
    results = []
    for cls in (wgth.GraphStructure, ref_mod.GraphStructure):
        g = cls()
        g.csr_row_ptr = g.csr_col_ind = holder
        results.append(g.multilayer_sample_without_replacement(seeds, fanout))
    ours, ref = results
    for a, b in zip(ours, ref):                      # target_gids, edge_indice, csr_row_ptr, csr_col_ind
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)
    per_class = len(calls) // 2
    assert [c[:2] for c in calls[:per_class]] == [c[:2] for c in calls[per_class:]]  # same centers and fan-out at every hop
