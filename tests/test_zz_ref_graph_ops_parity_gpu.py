"""Reference-BINARY parity of graph_append_unique / csr_add_self_loop on the GPU box: the reference's own kernels
(cpp/src/graph_ops/append_unique_func.cuh:47-353, csr_add_self_loop_func.cuh:24-58, built from /root/reference into
oracle/_ref/libwholegraph_ref.so and called through the same ctypes binding) vs this repo's.

append_unique: the reference appends new ids in hash-slot order (unspecified), this repo in first-occurrence order
(DESIGN.md deviation 8), so the comparison is what the reference's own Python test checks
(tests/wholegraph_torch/ops/test_graph_append_unique.py): targets stay in front, the appended parts are equal as SETS,
and each library's neighbor->unique mapping is consistent with its own list.  csr_add_self_loop: byte-exact.

(File name sorts last on purpose: the reference-side harness was added without a GPU at hand.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")


def _has_graph_ops():
    if not os.path.exists(REF_SO):
        return False
    try:
        out = subprocess.run(["nm", "-D", REF_SO], capture_output=True, text=True, timeout=60).stdout
        return " T graph_append_unique" in out and " T csr_add_self_loop" in out
    except Exception:
        return False


def _run_worker(tmp_path, name, lib=None):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    env.pop("WHOLEGRAPH_B200_LIB", None)
    if lib:
        env["WHOLEGRAPH_B200_LIB"] = lib
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_graph_ops_worker.py"), out], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, "worker failed:\n" + p.stdout[-2000:] + p.stderr[-4000:]
    return np.load(out)


@pytest.mark.skipif(not _has_graph_ops(), reason="oracle/_ref/libwholegraph_ref.so with graph_ops is not built "
                                                 "(needs /root/reference at build time)")
def test_reference_graph_ops_match_ours(tmp_path):
    import ref_graph_ops_worker as W
    ref = _run_worker(tmp_path, "ref", REF_SO)
    ours = _run_worker(tmp_path, "ours")
    for ci, (t, n, dt) in enumerate(W.UNIQUE_CASES):
        targets, neighbors = W.unique_inputs(ci)
        lists = {}
        for name, res in (("reference", ref), ("ours", ours)):
            uniq, mapping = res["unique_%d" % ci], res["mapping_%d" % ci]
            assert np.array_equal(uniq[:t], targets), (name, ci)
            assert len(np.unique(uniq)) == len(uniq), (name, ci, "duplicates in the unique list")
            assert mapping.shape[0] == n and (n == 0 or np.array_equal(uniq[mapping], neighbors)), (name, ci)
            lists[name] = uniq
        assert len(lists["ours"]) == len(lists["reference"]), ci
        assert np.array_equal(np.sort(lists["ours"][t:]), np.sort(lists["reference"][t:])), ci
    for ci in range(len(W.LOOP_CASES)):
        for k in ("loop_row_%d" % ci, "loop_col_%d" % ci):
            assert ref[k].tobytes() == ours[k].tobytes(), k
