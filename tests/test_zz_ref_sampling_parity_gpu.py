"""Reference-BINARY parity of the unweighted CSR sampler on the GPU box.

The reference's own sampling kernels (count -> exclusive scan -> per-node BlockRadixSort + pointer-jumping partial
Fisher-Yates, sample-all for k <= 0; cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh:39-475) are
built from /root/reference into oracle/_ref/libwholegraph_ref.so on top of a RESTATED PCG generator (RAFT is not
vendored; oracle/ref_shim/raft/random/rng_device.cuh, checked on CPU against the oracle stream and the pcg32 known-answer
vector in tests/test_ref_shim_rng.py).  So this pins the SELECTION algorithm, launch-shape table and output layout of this
repo's warp-per-node sampler against the reference's kernels, sample for sample; the random stream is the same restated
one on both sides and stays "unpinned against RAFT".  A third leg checks the oracle against the reference binary.

(File name sorts last on purpose: the reference-side harness was added without a GPU at hand.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")


def _has_sampler():
    if not os.path.exists(REF_SO):
        return False
    try:
        out = subprocess.run(["nm", "-D", REF_SO], capture_output=True, text=True, timeout=60).stdout
        return " T wholegraph_csr_unweighted_sample_without_replacement" in out
    except Exception:
        return False


def _run_worker(tmp_path, name, lib=None):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    env.pop("WHOLEGRAPH_B200_LIB", None)
    if lib:
        env["WHOLEGRAPH_B200_LIB"] = lib
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_sample_worker.py"), out], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, "worker failed:\n" + p.stdout[-2000:] + p.stderr[-4000:]
    return np.load(out)


def _oracle_results():
    import ref_sample_worker as W
    from oracle import oracle as O
    res = {}
    graphs = {dt: W.graph(dt) for dt in (np.int32, np.int64)}
    for ci, (k, cdt, ndt, n, seed) in enumerate(W.CASES):
        row_ptr, col = graphs[cdt]
        offs, dst, lid, gid = O.unweighted_sample(row_ptr, col, W.centers_of(ci), k, seed)
        for name, a in zip(("offsets", "dst", "center_lid", "edge_gid"), (offs, dst, lid, gid)):
            res["case%d_%s" % (ci, name)] = np.asarray(a)
    return res


@pytest.mark.skipif(not _has_sampler(), reason="oracle/_ref/libwholegraph_ref.so without the reference's sampler "
                                               "(needs /root/reference at build time)")
def test_reference_sampler_matches_ours_and_the_oracle(tmp_path):
    import ref_sample_worker as W
    ref = _run_worker(tmp_path, "ref", REF_SO)
    ours = _run_worker(tmp_path, "ours")
    exp = _oracle_results()
    assert sorted(ref.files) == sorted(ours.files) == sorted(exp)
    for ci, (k, cdt, ndt, n, seed) in enumerate(W.CASES):
        for name in ("offsets", "dst", "center_lid", "edge_gid"):
            key = "case%d_%s" % (ci, name)
            assert ref[key].tolist() == ours[key].tolist(), "ours differs from the reference binary: %s (k=%d)" % (key, k)
            assert ref[key].tolist() == exp[key].tolist(), "oracle differs from the reference binary: %s (k=%d)" % (key, k)
