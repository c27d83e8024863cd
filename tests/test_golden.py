"""CPU: the oracle must reproduce, byte for byte, what the REFERENCE's own kernels produced on a B200
(tests/golden/reference_gather_scatter_golden.npz, made by tools/make_golden.sh from oracle/_ref)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_reference_golden_vectors(monkeypatch):
    monkeypatch.setenv("WG_GOLDEN_SMALL", "1")
    sys.modules.pop("ref_parity_worker", None)
    sys.path.insert(0, HERE)
    import ref_parity_worker as W
    from oracle import oracle as O
    gold = np.load(os.path.join(HERE, "golden", "reference_gather_scatter_golden.npz"))
    assert W.ROWS == 307 and len(W.CASES) == 13
    checked = 0
    for ci, (mt, loc, tab_dt, out_dt, cols, stride, idt, n) in enumerate(W.CASES):
        table, idx, sentinel, sidx, src = W.case_inputs(ci)
        exp = sentinel.copy()
        O.gather(table, tab_dt, idx, out_dt, out=exp, cols=cols)
        assert np.array_equal(np.frombuffer(exp.tobytes(), dtype=np.uint8), gold["gather_%d" % ci]), "gather case %d" % ci
        tab = table.copy()
        O.scatter(src, out_dt, sidx, tab, tab_dt, cols=cols)
        assert np.array_equal(np.frombuffer(tab.tobytes(), dtype=np.uint8), gold["scatter_%d" % ci]), "scatter case %d" % ci
        checked += 2
    assert checked == 26
    sys.modules.pop("ref_parity_worker", None)
