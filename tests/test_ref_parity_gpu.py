"""Three-way bit-exact parity on the GPU box: the REFERENCE's own gather/scatter kernels (rebuilt from
/root/reference into oracle/_ref/libwholegraph_ref.so, loaded through the same ctypes binding) vs this repo's
sm_100a kernels vs the C oracle, on identical seeded inputs (13 cases: conversions, odd widths, strides,
HOST memory incl. the reference's sorted-ids branch).  This is what pins the oracle to the reference binary."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")


def _run_worker(tmp_path, name, lib=None):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    env.pop("WHOLEGRAPH_B200_LIB", None)
    if lib:
        env["WHOLEGRAPH_B200_LIB"] = lib
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_parity_worker.py"), out], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, "worker failed:\n" + p.stdout[-2000:] + p.stderr[-4000:]
    return np.load(out)


def _oracle_results():
    import ref_parity_worker as W
    from oracle import oracle as O
    res = {}
    for ci, (mt, loc, tab_dt, out_dt, cols, stride, idt, n) in enumerate(W.CASES):
        table, idx, sentinel, sidx, src = W.case_inputs(ci)
        exp = sentinel.copy()
        O.gather(table, tab_dt, idx, out_dt, out=exp, cols=cols)
        res["gather_%d" % ci] = np.frombuffer(exp.tobytes(), dtype=np.uint8)
        tab = table.copy()
        O.scatter(src, out_dt, sidx, tab, tab_dt, cols=cols)
        res["scatter_%d" % ci] = np.frombuffer(tab.tobytes(), dtype=np.uint8)
    return res


def test_ours_matches_oracle_in_worker(tmp_path):
    ours = _run_worker(tmp_path, "ours")
    exp = _oracle_results()
    for k, v in exp.items():
        assert np.array_equal(ours[k], v), k


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libwholegraph_ref.so not built (needs /root/reference at build time)")
def test_reference_binary_matches_oracle_and_ours(tmp_path):
    ref = _run_worker(tmp_path, "ref", REF_SO)
    ours = _run_worker(tmp_path, "ours")
    exp = _oracle_results()
    bad = [k for k in exp if not np.array_equal(ref[k], exp[k])]
    assert bad == [], "oracle differs from the reference binary on: %s" % bad
    bad = [k for k in exp if not np.array_equal(ref[k], ours[k])]
    assert bad == [], "this repo's kernels differ from the reference binary on: %s" % bad
