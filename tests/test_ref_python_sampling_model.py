"""A third, independent pin of the oracle's samplers: the HOST MODELS inside the reference's own Python tests
(tests/wholegraph_torch/ops/test_wholegraph_unweighted_sample_without_replacement.py:45-220 and
..._weighted_...:33-176), which its GPU outputs must match there.  Those models are plain Python over two host functions
of the library (generate_random_positive_int_cpu / generate_exponential_distribution_negative_float_cpu -- here THIS
repo's libwholegraph.so, CPU code).  The reference test modules are loaded unchanged from the reference tree through the
compat/ import alias and their models are run on CPU against the oracle:

  reference Python model (its launch-shape table, its Fisher-Yates base, its key formula)  x  this library's host RNG
      ==  oracle (C restatement of the reference's device kernels)

CPU only; skipped where /root/reference is absent."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPS = "/root/reference/python/pylibwholegraph/pylibwholegraph/tests/wholegraph_torch/ops"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_OPS), reason="reference tree not present")


def _load(name):
    for p in (ROOT, os.path.join(ROOT, "compat")):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF_OPS, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


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
