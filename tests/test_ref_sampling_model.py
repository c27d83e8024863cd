"""The oracle's neighbor samplers against the reference's OWN CPU models -- the host functions its GPU tests compare the
sampling kernels with (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu:
wholegraph_csr_unweighted_sample_without_replacement_cpu :419-505, wholegraph_csr_weighted_sample_without_replacement_cpu
:676-763), compiled for the CPU from /root/reference as they are (oracle/build_ref_host_sampling_model.sh) on top of the
restated PCG stand-in.  This is the pin SURVEY section 8(c) asks for: the oracle restates those functions, here it is
checked against them as CODE over fan-outs on both sides of every launch-shape boundary ((k-1)/32), k <= 0, int32 / int64
ids, float / double weights.  The random stream underneath stays the restated one (RAFT is not vendored).  CPU only.

Unweighted: all four outputs equal element for element.  Weighted: the reference's tests sort each center's samples before
comparing (segment_sort_output, :785-813), i.e. its contract is the SET per center; compared the same way here, and centers
whose k-th and (k+1)-th keys are within 1e-5 relative (the oracle's margin: a last-ulp tie between libm and the reference's
float expression) are exempt."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "ref_host_sampling_model.so")

pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/ref_host_sampling_model.so not built (needs /root/reference at build time)")

NODES = 1500


@pytest.fixture(scope="module")
def model(wmb):  # wmb: loads libwholegraph.so first (the model's descriptor helpers resolve to it)
    lib = ctypes.CDLL(SO, mode=os.RTLD_LAZY)
    vp, i64, c_int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    out = ctypes.POINTER(vp)
    lib.wgref_cpu_unweighted_sample.restype = c_int
    lib.wgref_cpu_unweighted_sample.argtypes = [vp, i64, vp, c_int, i64, vp, c_int, i64, c_int, ctypes.c_ulonglong, out, out, out, out]
    lib.wgref_cpu_weighted_sample.restype = c_int
    lib.wgref_cpu_weighted_sample.argtypes = [vp, i64, vp, c_int, i64, vp, c_int, vp, c_int, i64, c_int, ctypes.c_ulonglong, out, out, out, out]
    lib.wgref_model_free.argtypes = [vp]
    return lib


def _graph(col_dtype, seed=5):
    rng = np.random.default_rng(seed)
    deg = np.minimum((rng.pareto(1.1, size=NODES) * 8).astype(np.int64), 700)
    deg[::13] = 0
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, NODES, size=int(row_ptr[-1])).astype(col_dtype)
    return row_ptr, col


def _take(lib, ptr, dtype, count):
    a = np.frombuffer((ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr.value), dtype=dtype).copy() if count and ptr.value \
        else np.zeros(0, dtype)
    if ptr.value:
        lib.wgref_model_free(ptr)
    return a


def _reference_unweighted(lib, row_ptr, col, centers, k, seed):
    o, d, l, g = (ctypes.c_void_p() for _ in range(4))
    total = lib.wgref_cpu_unweighted_sample(row_ptr.ctypes.data, row_ptr.size, col.ctypes.data, int(col.dtype == np.int64), col.size,
                                            centers.ctypes.data, int(centers.dtype == np.int64), centers.size, k, seed,
                                            ctypes.byref(o), ctypes.byref(d), ctypes.byref(l), ctypes.byref(g))
    return (_take(lib, o, np.int32, centers.size + 1), _take(lib, d, col.dtype, total), _take(lib, l, np.int32, total),
            _take(lib, g, np.int64, total))


@pytest.mark.parametrize("k", [1, 5, 25, 31, 32, 33, 64, 65, 96, 97, 128, 129, 192, 193, 256, 257, 384, 385, 500, 1024, -1, 0])
@pytest.mark.parametrize("col_dtype,center_dtype", [(np.int32, np.int32), (np.int64, np.int64), (np.int64, np.int32)])
def test_unweighted_oracle_equals_the_reference_cpu_model(model, oracle, k, col_dtype, center_dtype):
    row_ptr, col = _graph(col_dtype)
    centers = np.random.default_rng(k + 1000).integers(0, NODES, size=257).astype(center_dtype)
    seed = 1234567 + k
    ref = _reference_unweighted(model, row_ptr, col, centers, k, seed)
    ours = oracle.unweighted_sample(row_ptr, col, centers, k, seed)
    for name, a, b in zip(("offsets", "dst", "center_local_id", "edge_gid"), ours, ref):
        assert a.tolist() == b.tolist(), "%s differs at k=%d" % (name, k)
    assert ref[0][-1] > 0


@pytest.mark.parametrize("k", [1, 5, 25, 32, 33, 100, 256, 257, 300])
@pytest.mark.parametrize("weight_dtype", [np.float32, np.float64])
def test_weighted_oracle_equals_the_reference_cpu_model(model, oracle, k, weight_dtype):
    row_ptr, col = _graph(np.int64, seed=9)
    rng = np.random.default_rng(k + 77)
    weights = (rng.random(col.size) * 4 + 0.05).astype(weight_dtype)
    centers = rng.integers(0, NODES, size=200).astype(np.int64)
    seed = 987 + k
    o, d, l, g = (ctypes.c_void_p() for _ in range(4))
    total = model.wgref_cpu_weighted_sample(row_ptr.ctypes.data, row_ptr.size, col.ctypes.data, 1, col.size, weights.ctypes.data,
                                            int(weight_dtype == np.float64), centers.ctypes.data, 1, centers.size, k, seed,
                                            ctypes.byref(o), ctypes.byref(d), ctypes.byref(l), ctypes.byref(g))
    r_off, r_dst = _take(model, o, np.int32, centers.size + 1), _take(model, d, np.int64, total)
    r_lid, r_gid = _take(model, l, np.int32, total), _take(model, g, np.int64, total)
    offs, dst, lid, gid, margin = oracle.weighted_sample(row_ptr, col, weights, centers, k, seed)
    assert offs.tolist() == r_off.tolist()
    assert lid.tolist() == r_lid.tolist()
    compared = 0
    for c in range(centers.size):
        b, e = int(offs[c]), int(offs[c + 1])
        if margin[c] < 1e-5:
            continue
        compared += 1
        assert sorted(gid[b:e].tolist()) == sorted(r_gid[b:e].tolist()), "edge set of center %d differs at k=%d" % (c, k)
        assert sorted(dst[b:e].tolist()) == sorted(r_dst[b:e].tolist())
    assert compared >= centers.size * 0.9
