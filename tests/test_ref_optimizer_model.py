"""The oracle's dedup + sparse optimizers against the reference's OWN CPU model of them, as code.

The reference's gradient-apply test carries a CPU model (class CPUOptimizer,
cpp/tests/wholememory_ops/wholememory_embedding_gradient_apply_tests.cu:169-371) which its GPU kernels must match at 1e-5
(same file, the comparison in TEST_P).  oracle/build_ref_host_optimizer_model.sh compiles that class for the CPU from where it
lies in the reference tree (into the git-ignored oracle/_ref/); here this repo's oracle -- the restatement of the
reference's device kernels that every GPU parity test is checked against -- is run through the same multi-step schedules
with duplicate ids and compared at that 1e-5 (and the fraction of bit-identical values is printed).  CPU only."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_host_optimizer_model.so")

pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/ref_host_optimizer_model.so not built "
                                "(needs /root/reference; built by __graft_entry__.build())")

KINDS = {"sgd": 1, "adam": 2, "rmsprop": 3, "adagrad": 4}


def _model():
    L = ctypes.CDLL(SO)
    L.wgref_cpu_optimizer_run.restype = ctypes.c_int
    L.wgref_cpu_optimizer_run.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def _run_model(kind, table, params, lr, steps_ids, steps_grads):
    L = _model()
    names = (ctypes.c_char_p * len(params))(*[k.encode() for k in params])
    values = np.array(list(params.values()), np.float32)
    counts = np.array([len(i) for i in steps_ids], np.int64)
    ids = np.ascontiguousarray(np.concatenate(steps_ids), np.int64)
    grads = np.ascontiguousarray(np.concatenate(steps_grads), np.float32)
    out = np.ascontiguousarray(table.copy(), np.float32)
    rc = L.wgref_cpu_optimizer_run(KINDS[kind], out.shape[0], out.shape[1], names, values.ctypes.data, len(params), lr,
                                   len(steps_ids), counts.ctypes.data, ids.ctypes.data, grads.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def _run_oracle(kind, table, params, lr, steps_ids, steps_grads):
    w = table.copy()
    rows = w.shape[0]
    m, v, b12 = np.zeros_like(w), np.zeros_like(w), np.ones((rows, 2), np.float32)
    kw = {k: params[k] for k in ("weight_decay", "epsilon", "beta1", "beta2", "alpha") if k in params}
    for ids, g in zip(steps_ids, steps_grads):
        urows, ug = O.dedup_gradients(ids, g)
        if kind == "adam":
            O.optimizer_step("adam", w, urows, ug, lr, state=(m, v), b12=b12, adam_w=params.get("adam_w", 0) > 0.5, **kw)
        elif kind == "sgd":
            O.optimizer_step("sgd", w, urows, ug, lr, weight_decay=kw.get("weight_decay", 0.0))
        else:
            O.optimizer_step(kind, w, urows, ug, lr, state=m, **kw)
    return w


CASES = [
    ("sgd", {}),
    ("sgd", {"weight_decay": 0.05}),
    ("adam", {}),
    ("adam", {"weight_decay": 0.01, "beta1": 0.85, "beta2": 0.98, "epsilon": 1e-6}),
    ("adam", {"weight_decay": 0.02, "adam_w": 1.0}),
    ("rmsprop", {}),
    ("rmsprop", {"weight_decay": 0.01, "alpha": 0.9, "epsilon": 1e-6}),
    ("adagrad", {}),
    ("adagrad", {"weight_decay": 0.03, "epsilon": 1e-7}),
]


@pytest.mark.parametrize("kind,params", CASES, ids=[f"{k}-{'-'.join(p) or 'defaults'}" for k, p in CASES])
@pytest.mark.parametrize("rows,dim,per_step,steps", [(97, 8, 40, 6), (400, 33, 700, 4), (1000, 128, 300, 10)])
def test_oracle_matches_the_reference_cpu_optimizer(kind, params, rows, dim, per_step, steps):
    rng = np.random.default_rng(hash((kind, rows, dim, len(params))) & 0xFFFF)
    table = rng.standard_normal((rows, dim)).astype(np.float32)
    steps_ids = [rng.integers(0, rows, per_step).astype(np.int64) for _ in range(steps)]  # duplicates by construction
    steps_grads = [rng.standard_normal((per_step, dim)).astype(np.float32) for _ in range(steps)]
    lr = 0.01
    want = _run_model(kind, table, params, lr, steps_ids, steps_grads)
    got = _run_oracle(kind, table, params, lr, steps_ids, steps_grads)
    exact = float(np.mean(want.view(np.uint32) == got.view(np.uint32)))
    print(f"{kind} {params}: {exact:.4f} of values bit-identical, max |diff| {np.abs(want - got).max():.3e}")
    # the tolerance the reference holds its own kernels to against this model
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    touched = np.unique(np.concatenate(steps_ids))
    untouched = np.setdiff1d(np.arange(rows), touched)
    assert np.all(np.any(want[touched] != table[touched], axis=1))  # every touched row really moved
    np.testing.assert_array_equal(got[untouched], table[untouched])
    np.testing.assert_array_equal(want[untouched], table[untouched])


def test_the_comparison_is_not_vacuous():
    """A different learning rate on one side must show up as a mismatch."""
    rng = np.random.default_rng(5)
    table = rng.standard_normal((50, 16)).astype(np.float32)
    ids = [rng.integers(0, 50, 64).astype(np.int64) for _ in range(3)]
    grads = [rng.standard_normal((64, 16)).astype(np.float32) for _ in range(3)]
    for kind in KINDS:
        want = _run_model(kind, table, {}, 0.01, ids, grads)
        got = _run_oracle(kind, table, {}, 0.0101, ids, grads)
        assert np.abs(want - got).max() > 1e-5, kind
