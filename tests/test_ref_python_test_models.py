"""A third, independent pin of the oracle's samplers: the HOST MODELS inside the reference's own Python tests
(tests/wholegraph_torch/ops/test_wholegraph_unweighted_sample_without_replacement.py:45-220 and
..._weighted_...:33-176), which its GPU outputs must match there.  Those models are plain Python over two host functions
of the library (generate_random_positive_int_cpu / generate_exponential_distribution_negative_float_cpu -- here THIS
repo's libwholegraph.so, CPU code).  The reference test modules are loaded unchanged from the reference tree through the
compat/ import alias and their models are run on CPU against the oracle:

  reference Python model (its launch-shape table, its Fisher-Yates base, its key formula)  x  this library's host RNG
      ==  oracle (C restatement of the reference's device kernels)

CPU only; skipped where /root/reference is absent."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPS = "/root/reference/python/pylibwholegraph/pylibwholegraph/tests/wholegraph_torch/ops"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_OPS), reason="reference tree not present")


def _load(name):
    from compat_loader import load_reference_file
    return load_reference_file(os.path.join(REF_OPS, name + ".py"), "_ref_" + name)


def _graph(seed, nodes, edges, col_dtype, weight_dtype=torch.float32):
    from wholegraph_b200.test_utils.test_comm import gen_csr_graph
    torch.manual_seed(seed)
    return gen_csr_graph(nodes, edges, csr_col_dtype=col_dtype, weight_dtype=weight_dtype)


@pytest.mark.parametrize("k", [5, 11, 33, 70, -1])           # launch-shape rows 0, 0, 1, 2 of the reference's table; take-all
@pytest.mark.parametrize("col_dtype", [torch.int32, torch.int64])
def test_reference_python_unweighted_model_equals_the_oracle(k, col_dtype):
    from oracle import oracle as O
    ref = _load("test_wholegraph_unweighted_sample_without_replacement")
    nodes, edges = 40, 40 * 39  # dense enough that most centers have more than k neighbours... degree 39
    if k == 70:
        nodes, edges = 90, 90 * 80
    row_ptr, col, _ = _graph(k + 100, nodes, edges, col_dtype)
    centers = torch.randint(0, nodes, (13,), dtype=torch.int64, generator=torch.Generator().manual_seed(k + 7))
    seed = 4321 + k
    want = ref.host_unweighted_sample_without_replacement(row_ptr, col, centers, k, col_dtype, seed)
    got = O.unweighted_sample(row_ptr.numpy(), col.numpy().astype(np.int64), centers.numpy(), k, seed)
    assert k <= 0 or int(want[0][-1]) == 13 * k            # every center really went through the sampling branch
    for name, w, g in zip(("offsets", "dst", "center_lid", "edge_gid"), want, got):
        assert torch.as_tensor(w).to(torch.int64).tolist() == np.asarray(g).astype(np.int64).tolist(), (name, k)


@pytest.mark.parametrize("k", [5, 11])
@pytest.mark.parametrize("weight_dtype", [torch.float32, torch.float64])
def test_reference_python_weighted_model_equals_the_oracle(k, weight_dtype):
    import wholegraph_b200.binding as wmb
    from oracle import oracle as O
    ref = _load("test_wholegraph_weighted_sample_without_replacement")
    nodes, edges = 30, 30 * 24
    row_ptr, col, weights = _graph(k, nodes, edges, torch.int32, weight_dtype)
    centers = torch.randint(0, nodes, (9,), dtype=torch.int64, generator=torch.Generator().manual_seed(k))
    seed = 99 + k
    want = ref.host_weighted_sample_without_replacement(row_ptr, col, weights, centers, k, wmb.WholeMemoryDataType.DtInt, seed)
    eo, ed, el, eg, margin = O.weighted_sample(row_ptr.numpy(), col.numpy().astype(np.int64), weights.numpy(), centers.numpy(), k, seed)
    assert want[0].tolist() == eo.tolist() and int(eo[-1]) == 9 * k
    assert want[2].tolist() == el.tolist()
    w_gid, loose = want[3].numpy(), 0
    for c in range(centers.shape[0]):                       # per-center sets: the reference test's own comparison
        a, b = eo[c], eo[c + 1]
        if sorted(w_gid[a:b].tolist()) != sorted(eg[a:b].tolist()):
            # the Python model draws its keys with the HOST formula (double log1p), the oracle with the device formula
            # (float log1pf): only a key pair closer than the oracle's margin may resolve differently
            assert margin[c] < 1e-5, (c, margin[c])
            loose += 1
    assert loose <= 1
    assert np.array_equal(col.numpy()[w_gid].astype(np.int64), want[1].numpy().astype(np.int64))


def test_reference_python_graph_op_models_equal_the_expectations_of_this_repos_gpu_tests():
    """The host models of the reference's append_unique / add_csr_self_loop Python tests against the expectations this repo's
    GPU tests hold the kernels to (tests/test_graph_ops_gpu.py: first-occurrence order, closed-form self-loop CSR)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_graph_ops_gpu as ours
    loop = _load("test_graph_add_csr_self_loop")
    uniq = _load("test_graph_append_unique")
    for seed, (rows, nbr, edges) in enumerate([(101, 157, 1001), (113, 1987, 2305), (5, 9, 0)]):
        row_ptr, col, _ = _graph(seed, rows, edges, torch.int32)
        if nbr != rows:
            torch.manual_seed(seed)
            from wholegraph_b200.test_utils.test_comm import gen_csr_graph
            row_ptr, col, _ = gen_csr_graph(rows, edges, nbr, csr_row_dtype=torch.int32, csr_col_dtype=torch.int32)
        want_row, want_col = loop.host_add_csr_self_loop(row_ptr.to(torch.int32), col)
        rp = row_ptr.numpy().astype(np.int64)
        exp_row = rp + np.arange(rows + 1)
        assert want_row.numpy().tolist() == exp_row.tolist()
        for r in range(rows):
            seg = want_col.numpy()[exp_row[r]:exp_row[r + 1]]
            assert seg[0] == r and np.array_equal(seg[1:], col.numpy()[rp[r]:rp[r + 1]])
    g = torch.Generator().manual_seed(3)
    for t, n, dt in ((10, 104, torch.int32), (113, 1987, torch.int64)):
        targets = torch.randperm(n, generator=g, dtype=dt)[:t]
        neighbors = torch.randint(0, n, (n,), generator=g, dtype=dt)
        exp_uniq, exp_map = ours._first_occurrence_reference(targets.tolist(), neighbors.tolist())
        # the reference test accepts any order of the appended part: set equality + a mapping consistent with the list
        assert sorted(exp_uniq) == torch.unique(torch.cat((targets, neighbors))).tolist()
        want_map = uniq.host_neighbor_raw_to_unique(torch.tensor(exp_uniq, dtype=dt), neighbors)
        assert want_map.dtype == torch.int32 and want_map.tolist() == exp_map
