"""The helper modules the reference's Python tests import (`pylibwholegraph.utils.multiprocess`,
`pylibwholegraph.test_utils.test_comm`) exist here under the same names, give the same results as the reference's own
helpers, and with them EVERY Python test module of the reference imports unchanged against this implementation.
CPU only; the comparison needs the reference tree (skipped where /root/reference is absent)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = "/root/reference/python/pylibwholegraph/pylibwholegraph"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_PY), reason="reference tree not present")


def _helper_pair():
    """(this repo's test_comm, the reference's test_comm loaded from its file with its imports resolved by compat/)"""
    from compat_loader import activate_compat, load_reference_file
    activate_compat()
    import pylibwholegraph.test_utils.test_comm as ours
    return ours, load_reference_file(os.path.join(REF_PY, "test_utils", "test_comm.py"), "_reference_test_comm")


@needs_reference
def test_helpers_give_the_reference_helpers_results():
    ours, ref = _helper_pair()
    for total, world in ((1000, 4), (1024 * 256 * 8 + 3, 8), (17, 3), (5, 1)):
        a, b = ours.random_partition(total, world), ref.random_partition(total, world)
        assert a.dtype == b.dtype and np.array_equal(a, b) and int(a.sum()) == total
    for fn in ("int_to_wholememory_datatype", "int_to_wholememory_location", "int_to_wholememory_type"):
        for v in range(-1, 5):
            try:
                want = getattr(ref, fn)(v)
            except ValueError:
                with pytest.raises(ValueError):
                    getattr(ours, fn)(v)
                continue
            if want is None:  # the reference falls through for negative values of some maps
                continue
            assert getattr(ours, fn)(v) == want, (fn, v)
    for seed, (nodes, edges, nbr, col_dt, w_dt) in enumerate([(103, 1043, None, torch.int32, torch.float32), (113, 1043, None, torch.int64, torch.float64),
                                                              (40, 300, 57, torch.int32, torch.float32), (7, 49, None, torch.int64, torch.float32),
                                                              (9, 0, None, torch.int32, torch.float32)]):
        torch.manual_seed(seed)
        a = ours.gen_csr_graph(nodes, edges, nbr, csr_col_dtype=col_dt, weight_dtype=w_dt)
        torch.manual_seed(seed)
        b = ref.gen_csr_graph(nodes, edges, nbr, csr_col_dtype=col_dt, weight_dtype=w_dt)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and torch.equal(x, y)
        row_ptr, col, _w = a
        for cen_dt in (torch.int32, torch.int64):
            centers = torch.randint(0, nodes, (13,), dtype=cen_dt)
            for k in (11, -1, 3):
                oa, ob = ours.host_get_sample_offset_tensor(row_ptr, centers, k), ref.host_get_sample_offset_tensor(row_ptr, centers, k)
                assert oa.dtype == ob.dtype and torch.equal(oa, ob)
            full = ours.host_get_sample_offset_tensor(row_ptr, centers, -1)
            total = int(full[-1])
            ra = ours.host_sample_all_neighbors(row_ptr, col, centers, full, col.dtype, total)
            rb = ref.host_sample_all_neighbors(row_ptr, col, centers, full.clone(), col.dtype, total)
            for x, y in zip(ra, rb):
                assert x.dtype == y.dtype and torch.equal(x, y)


def _two_rank_body(rank, world):
    assert world == 2 and rank in (0, 1)
    if os.environ.get("WG_TEST_FAIL_RANK") == str(rank):
        raise SystemExit(3)


def test_multiprocess_run_spawns_ranks_and_reports_failures(monkeypatch):
    from wholegraph_b200.utils.multiprocess import multiprocess_run
    seen = []
    multiprocess_run(1, lambda r, w: seen.append((r, w)), inline_single_process=True)
    assert seen == [(0, 1)]
    multiprocess_run(2, _two_rank_body)
    monkeypatch.setenv("WG_TEST_FAIL_RANK", "1")
    with pytest.raises(AssertionError):
        multiprocess_run(2, _two_rank_body)


@needs_reference
def test_every_reference_python_test_module_imports_against_this_implementation():
    files = sorted(glob.glob(REF_PY + "/tests/pylibwholegraph/test_*.py") + glob.glob(REF_PY + "/tests/wholegraph_torch/ops/test_*.py"))
    assert len(files) >= 9
    program = ("import importlib.util, sys\n"
               "import wholegraph_b200.binding\n"
               "for f in sys.argv[1:]:\n"
               "    spec = importlib.util.spec_from_file_location('_ref_test_module', f)\n"
               "    m = importlib.util.module_from_spec(spec)\n"
               "    spec.loader.exec_module(m)\n"
               "    tests = [n for n in dir(m) if n.startswith('test_')]\n"
               "    assert tests, f\n"
               "    print('imported', f.rsplit('/', 1)[1], len(tests))\n"
               "import pylibwholegraph.torch, wholegraph_b200.torch\n"
               "assert pylibwholegraph.torch is wholegraph_b200.torch\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat")]))
    p = subprocess.run([sys.executable, "-c", program] + files, capture_output=True, text=True, timeout=600, env=env, cwd="/")
    assert p.returncode == 0 and p.stdout.count("imported") == len(files), p.stdout + p.stderr[-3000:]
