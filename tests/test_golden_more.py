"""CPU: the oracle against golden outputs of the REFERENCE's own optimizer / sampler / graph-op kernels
(tests/golden/reference_{optimizer,sampler,graph_ops}_golden.npz, produced on a B200 by tools/make_golden.sh from
oracle/_ref/libwholegraph_ref.so running the seeded case lists of tests/ref_*_worker.py).

The fixtures need one GPU run of the reference binary; until they are committed these tests skip (the live three-way
comparison is tests/test_zz_ref_*_parity_gpu.py).  Inputs are regenerated from the workers' seeds, like test_golden.py."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(HERE, "golden")


WORKERS = ("ref_sample_worker", "ref_optimizer_worker", "ref_graph_ops_worker", "test_zz_ref_optimizer_parity_gpu")


@pytest.fixture(autouse=True)
def _compact_case_lists(monkeypatch):
    """The fixtures were made with WG_GOLDEN_SMALL=1 (compact sizes); the workers read it at import."""
    monkeypatch.setenv("WG_GOLDEN_SMALL", "1")
    for m in WORKERS:
        sys.modules.pop(m, None)
    yield
    for m in WORKERS:
        sys.modules.pop(m, None)


def _gold(name):
    path = os.path.join(GOLD, "reference_%s_golden.npz" % name)
    if not os.path.exists(path):
        pytest.skip("tests/golden/reference_%s_golden.npz not generated yet (tools/make_golden.sh on a GPU box)" % name)
    return np.load(path)


def test_oracle_sampler_matches_reference_kernel_outputs():
    gold = _gold("sampler")
    import ref_sample_worker as W
    from oracle import oracle as O
    graphs = {dt: W.graph(dt) for dt in (np.int32, np.int64)}
    for ci, (k, cdt, ndt, n, seed) in enumerate(W.CASES):
        row_ptr, col = graphs[cdt]
        got = O.unweighted_sample(row_ptr, col, W.centers_of(ci), k, seed)
        for name, a in zip(("offsets", "dst", "center_lid", "edge_gid"), got):
            assert np.asarray(a).tolist() == gold["case%d_%s" % (ci, name)].tolist(), (ci, name, k)


def test_oracle_optimizers_match_reference_kernel_outputs():
    gold = _gold("optimizer")
    import test_zz_ref_optimizer_parity_gpu as T
    exp = T._oracle_results()
    assert sorted(exp) == sorted(gold.files)
    T._compare({k: gold[k] for k in gold.files}, exp, "oracle (CPU restatement) vs the reference binary's golden outputs")


def test_numpy_restatement_of_graph_ops_matches_reference_kernel_outputs():
    """No C oracle for the two graph ops: their contract is restated in numpy here (set semantics for append_unique -- the
    reference leaves the order of the appended part unspecified -- and the closed form of the self-loop CSR)."""
    gold = _gold("graph_ops")
    import ref_graph_ops_worker as W
    for ci, (t, n, dt) in enumerate(W.UNIQUE_CASES):
        targets, neighbors = W.unique_inputs(ci)
        uniq, mapping = gold["unique_%d" % ci], gold["mapping_%d" % ci]
        assert np.array_equal(uniq[:t], targets)
        assert np.array_equal(np.sort(uniq), np.unique(np.concatenate([targets, neighbors])))
        assert mapping.shape[0] == n and (n == 0 or np.array_equal(uniq[mapping], neighbors))
    for ci, (rows, _max_deg) in enumerate(W.LOOP_CASES):
        row_ptr, col = W.loop_inputs(ci)
        exp_row = row_ptr + np.arange(rows + 1, dtype=np.int32)
        exp_col = np.empty(col.size + rows, dtype=np.int32)
        for r in range(rows):  # self edge first, then the row's edges in their order
            exp_col[exp_row[r]] = r
            exp_col[exp_row[r] + 1:exp_row[r + 1]] = col[row_ptr[r]:row_ptr[r + 1]]
        assert np.array_equal(gold["loop_row_%d" % ci], exp_row) and np.array_equal(gold["loop_col_%d" % ci], exp_col)
