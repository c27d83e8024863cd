"""GPU tests of graph_append_unique / csr_add_self_loop, modelled on the reference's Python tests
(python/pylibwholegraph/pylibwholegraph/tests/wholegraph_torch/ops/test_graph_append_unique.py, test_graph_add_csr_self_loop.py):
set equality with torch.unique, targets kept in front, mapping consistent with the returned list -- plus the stronger
property this implementation guarantees (appended part in order of first occurrence)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wgth():
    import gpu_utils as G
    G.single_comm()
    import wholegraph_b200.torch as t
    return t


def _first_occurrence_reference(targets, neighbors):
    seen = {int(v): i for i, v in enumerate(targets)}
    uniq = [int(v) for v in targets]
    mapping = []
    for v in neighbors:
        v = int(v)
        if v not in seen:
            seen[v] = len(uniq)
            uniq.append(v)
        mapping.append(seen[v])
    return uniq, mapping


@pytest.mark.parametrize("target_count,neighbor_count", [(10, 100), (113, 1987), (0, 50), (64, 0), (20000, 200000)])
@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("need_map", [True, False])
def test_append_unique(wgth, target_count, neighbor_count, dtype, need_map):
    g = torch.Generator()
    g.manual_seed(target_count * 31 + neighbor_count)
    universe = max(neighbor_count, target_count, 1)
    targets = torch.randperm(universe, generator=g, dtype=dtype)[:target_count]
    neighbors = torch.randint(0, universe, (neighbor_count,), generator=g, dtype=dtype)
    res = wgth.append_unique(targets.cuda(), neighbors.cuda(), need_neighbor_raw_to_unique=need_map)
    uniq = (res[0] if need_map else res).cpu()
    ref_sorted = torch.unique(torch.cat((targets, neighbors)), sorted=True)
    assert torch.equal(torch.sort(uniq)[0], ref_sorted)           # reference test: set equality
    assert torch.equal(uniq[:target_count], targets)              # targets unchanged, in front
    exp_uniq, exp_map = _first_occurrence_reference(targets.tolist(), neighbors.tolist())
    assert uniq.tolist() == exp_uniq                              # deterministic first-occurrence order
    if need_map:
        m = res[1].cpu()
        assert m.dtype == torch.int32 and m.tolist() == exp_map
        if neighbor_count:
            assert torch.equal(uniq[m.long()], neighbors)         # reference test: mapping consistent with the list


def test_append_unique_docstring_example(wgth):
    t = torch.tensor([3, 11, 2, 10], dtype=torch.int64).cuda()
    n = torch.tensor([4, 5, 2, 11, 6, 9, 10, 5], dtype=torch.int64).cuda()
    u, m = wgth.append_unique(t, n, need_neighbor_raw_to_unique=True)
    assert u.tolist() == [3, 11, 2, 10, 4, 5, 6, 9] and m.tolist() == [4, 5, 2, 1, 6, 7, 3, 5]


@pytest.mark.parametrize("rows,max_deg", [(1, 5), (37, 9), (5000, 40)])
def test_add_csr_self_loop(wgth, rows, max_deg):
    rng = np.random.default_rng(rows)
    deg = rng.integers(0, max_deg, size=rows)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, rows, size=int(row_ptr[-1])).astype(np.int32)
    orp, oc = wgth.add_csr_self_loop(torch.from_numpy(row_ptr).cuda(), torch.from_numpy(col).cuda())
    orp, oc = orp.cpu().numpy(), oc.cpu().numpy()
    assert np.array_equal(orp, row_ptr + np.arange(rows + 1))
    for r in range(rows):
        seg = oc[orp[r]:orp[r + 1]]
        assert seg[0] == r and np.array_equal(seg[1:], col[row_ptr[r]:row_ptr[r + 1]])
