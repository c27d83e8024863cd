"""CPU tests that PIN the oracle (oracle/wm_oracle.c) before anything is checked against it.

Pins used (the reference ships no golden files, its tests are generator based):
  * the reference tests' closed-form table pattern (embedding_test_utils.cu:197-238): a gathered row
    is a pure function of its index, so expected outputs need no second implementation;
  * an independent numpy restatement of the conversion chain (numpy's own IEEE fp16 rounding);
  * exhaustive fp16 and dense bf16 conversion sweeps;
  * the published PCG32 known-answer vector (pcg32-demo, seed 42 / stream 54);
  * a line-for-line Python transcription of the reference's CPU sampler
    (graph_sampling_test_utils.cu:306-321).
"""
import itertools

import numpy as np
import pytest

from oracle import oracle as O


def _rand_table(rng, dt, rows, cols):
    if dt in O.INT_DTS:
        info = np.iinfo(O.NP_OF[dt])
        return rng.integers(info.min, info.max, size=(rows, cols), dtype=O.NP_OF[dt], endpoint=True)
    if dt == O.DT_BF16:
        return O._f32_to_bf16(rng.standard_normal((rows, cols)).astype(np.float32) * 100)
    if dt == O.DT_DOUBLE:
        # values that exercise double rounding on the way to fp16
        return rng.standard_normal((rows, cols)) * rng.choice([1e-6, 1.0, 1e3, 7e4], size=(rows, cols))
    return (rng.standard_normal((rows, cols)) * rng.choice([1e-6, 1.0, 1e3, 7e4], size=(rows, cols))).astype(O.NP_OF[dt])


@pytest.mark.parametrize("src,dst", list(itertools.product(O.FLOAT_DTS, O.FLOAT_DTS)) + list(itertools.product(O.INT_DTS, O.INT_DTS)))
def test_gather_c_matches_numpy_restatement(src, dst):
    rng = np.random.default_rng(1234 + src * 16 + dst)
    table = _rand_table(rng, src, 257, 19)
    idx = rng.integers(-3, 257, size=1000).astype(np.int64)
    got = O.gather(table, src, idx, dst)
    exp = O.np_gather(table, src, idx, dst)
    assert got.tobytes() == exp.tobytes()


@pytest.mark.parametrize("dt", O.FLOAT_DTS + O.INT_DTS)
@pytest.mark.parametrize("idx_dtype", [np.int32, np.int64])
def test_gather_closed_form_pattern(dt, idx_dtype):
    """Reference gather test: table row r == convert(r & mask) in every column, compare raw bits."""
    rows, cols, stride = 5000, 11, 12
    table = O.test_pattern(dt, 0, rows, cols, stride)
    rng = np.random.default_rng(7)
    idx = rng.integers(0, rows, size=3001).astype(idx_dtype)
    got = O.gather(table, dt, idx, dt, cols=cols)
    expected = O.test_pattern(dt, 0, rows, cols, cols)[idx.astype(np.int64)]
    assert got.tobytes() == expected.tobytes()
    # and the pattern itself: mask is 2^(M+1)-1 with M the mantissa bits
    mant = {O.DT_FLOAT: 23, O.DT_DOUBLE: 52, O.DT_HALF: 10, O.DT_BF16: 7}
    r = 4097
    if dt in mant:
        v = r & ((1 << (mant[dt] + 1)) - 1)
        as_f32 = O.convert(table[r:r + 1, :1], dt, O.DT_FLOAT)[0, 0] if dt != O.DT_DOUBLE else table[r, 0]
        assert float(as_f32) == float(v)
    else:
        assert int(table[r, 0]) == int(np.array(r).astype(O.NP_OF[dt]))


def test_fp16_conversion_exhaustive():
    allh = np.arange(65536, dtype=np.uint16).view(np.float16).reshape(1, -1)
    f32 = O.convert(allh, O.DT_HALF, O.DT_FLOAT)
    ref = allh.astype(np.float32)
    nan = np.isnan(ref)
    assert np.array_equal(f32.view(np.uint32)[~nan], ref.view(np.uint32)[~nan]) and np.all(np.isnan(f32[nan]))
    back = O.convert(f32, O.DT_FLOAT, O.DT_HALF)
    assert np.array_equal(back.view(np.uint16)[~nan], allh.view(np.uint16)[~nan])


def test_float_to_half_rounding_matches_ieee():
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2**32, size=2_000_00, dtype=np.uint64).astype(np.uint32)
    # add the interesting boundaries: ties, subnormals, overflow edge
    edge = np.array([0x33000000, 0x33000001, 0x337FFFFF, 0x38800000, 0x387FFFFF, 0x477FE000, 0x477FEFFF, 0x477FF000,
                     0x47800000, 0x3F801000, 0x3F803000, 0x7F800000, 0xFF800000, 0x00000001, 0x80000000], dtype=np.uint32)
    f = np.concatenate([bits, edge]).view(np.float32).reshape(1, -1)
    got = O.convert(f, O.DT_FLOAT, O.DT_HALF).view(np.uint16)
    with np.errstate(over="ignore"):
        ref = f.astype(np.float16).view(np.uint16)
    nan = np.isnan(f)
    assert np.array_equal(got[~nan], ref[~nan])
    assert np.all((got[nan] & 0x7C00) == 0x7C00) and np.all((got[nan] & 0x03FF) != 0)


def test_bf16_conversion_round_to_nearest_even():
    f = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.3895313892515355e38, 1e-40], dtype=np.float32).reshape(1, -1)
    got = O.convert(f, O.DT_FLOAT, O.DT_BF16)
    # 1.00390625 = 0x3F808000 is a tie -> even (0x3F80); 1.01171875 = 0x3F818000 tie -> even (0x3F82)
    assert got[0, 0] == 0x3F80 and got[0, 1] == 0x3F80 and got[0, 2] == 0x3F82 and got[0, 3] == 0xC020
    back = O.convert(got, O.DT_BF16, O.DT_FLOAT)
    assert back[0, 0] == 1.0 and back[0, 3] == -2.5


def test_double_to_half_rounds_twice_like_the_reference():
    # 1 + 2^-11 + 2^-30: a single rounding gives 1+2^-10, double rounding (via float32) gives 1.0
    x = np.array([[1.0 + 2.0**-11 + 2.0**-30]], dtype=np.float64)
    got = O.convert(x, O.DT_DOUBLE, O.DT_HALF).view(np.float16)[0, 0]
    assert float(got) == 1.0
    assert float(x.astype(np.float16)[0, 0]) == 1.0 + 2.0**-10  # what a "fixed" conversion would produce


def test_scatter_then_gather_roundtrip_and_negative_indices():
    rng = np.random.default_rng(11)
    table = np.zeros((100, 8), dtype=np.float32)
    src = rng.standard_normal((40, 8)).astype(np.float32)
    idx = rng.permutation(100)[:40].astype(np.int64)
    idx[5] = -1
    O.scatter(src, O.DT_FLOAT, idx, table, O.DT_FLOAT)
    out = np.full((40, 8), 7.0, dtype=np.float32)
    O.gather(table, O.DT_FLOAT, idx, O.DT_FLOAT, out=out)
    keep = idx >= 0
    assert np.array_equal(out[keep], src[keep])
    assert np.all(out[5] == 7.0)  # skipped row untouched
    assert not table.any(axis=1)[np.setdiff1d(np.arange(100), idx[keep])].any()


def test_partition_plan():
    # ceil(N/ws) per rank, tail ranks short or empty (memory_handle.cpp:1618-1635)
    assert O.partition(10, 4).tolist() == [0, 3, 6, 9, 10]
    assert O.partition(3, 8).tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    assert O.partition(1_000_000_000, 8).tolist() == [i * 125_000_000 for i in range(9)]
    assert O.partition(0, 2).tolist() == [0, 0, 0]


def _np_adam(w, m, v, b1t, b2t, g, lr, wd, eps, b1, b2, adam_w):
    f = np.float32
    b1t, b2t = f(b1t * f(b1)), f(b2t * f(b2))
    if adam_w:
        w = w - f(lr) * f(wd) * w
    else:
        g = g + f(wd) * w
    m = f(b1) * m + (f(1) - f(b1)) * g
    v = f(b2) * v + (f(1) - f(b2)) * g * g
    mhat = m / (f(1) - b1t)
    vhat = v / (f(1) - b2t)
    w = w - f(lr) * mhat / (np.sqrt(vhat) + f(eps))
    return w, m, v, b1t, b2t


@pytest.mark.parametrize("adam_w", [False, True])
def test_lazy_adam_against_numpy_float32(adam_w):
    rng = np.random.default_rng(5)
    N, D = 50, 13
    w = rng.standard_normal((N, 16)).astype(np.float32)
    m = np.zeros((N, 16), np.float32)
    v = np.zeros((N, 16), np.float32)
    b12 = np.ones((N, 2), np.float32)
    w0 = w.copy()
    ew, em, ev, eb = w.copy(), m.copy(), v.copy(), b12.copy()
    for step in range(3):
        rows = rng.permutation(N)[:20].astype(np.int64)
        g = rng.standard_normal((20, D)).astype(np.float32)
        O.optimizer_step("adam", w, rows, g, 0.01, state=(m, v), b12=b12, weight_decay=0.01, adam_w=adam_w, dim=D)
        for k, r in enumerate(rows):
            ew[r, :D], em[r, :D], ev[r, :D], eb[r, 0], eb[r, 1] = _np_adam(
                ew[r, :D], em[r, :D], ev[r, :D], eb[r, 0], eb[r, 1], g[k], 0.01, 0.01, 1e-8, 0.9, 0.999, adam_w)
    np.testing.assert_allclose(w, ew, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m, em, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(v, ev, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b12, eb, rtol=1e-6)
    assert np.array_equal(w[:, D:], w0[:, D:])  # padding columns never touched


def test_sgd_adagrad_rmsprop_against_numpy_float32():
    rng = np.random.default_rng(6)
    N, D, f = 30, 8, np.float32
    rows = np.arange(0, 30, 3).astype(np.int64)
    g = rng.standard_normal((rows.size, D)).astype(f)
    w = rng.standard_normal((N, D)).astype(f)
    e = w.copy()
    O.optimizer_step("sgd", w, rows, g, 0.1, weight_decay=0.05)
    gg = g + f(0.05) * e[rows]
    e[rows] = e[rows] - f(0.1) * gg
    np.testing.assert_allclose(w, e, rtol=1e-6, atol=1e-7)

    w = rng.standard_normal((N, D)).astype(f)
    s = np.abs(rng.standard_normal((N, D))).astype(f)
    e, es = w.copy(), s.copy()
    O.optimizer_step("adagrad", w, rows, g, 0.1, state=s, epsilon=1e-6)
    es[rows] = es[rows] + g * g
    e[rows] = e[rows] - f(0.1) * g / (np.sqrt(es[rows]) + f(1e-6))
    np.testing.assert_allclose(w, e, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(s, es, rtol=1e-6)

    w = rng.standard_normal((N, D)).astype(f)
    s = np.abs(rng.standard_normal((N, D))).astype(f)
    e, es = w.copy(), s.copy()
    O.optimizer_step("rmsprop", w, rows, g, 0.1, state=s, epsilon=1e-6, alpha=0.9)
    es[rows] = f(0.9) * es[rows] + (f(1) - f(0.9)) * g * g
    e[rows] = e[rows] - f(0.1) * g / (np.sqrt(es[rows]) + f(1e-6))
    np.testing.assert_allclose(w, e, rtol=1e-6, atol=1e-7)


def test_dedup_sums_in_arrival_order():
    ids = np.array([5, 2, 5, 9, 2, 5], dtype=np.int64)
    g = np.array([[1e8], [1.0], [-1e8], [3.0], [2.0], [1.0]], dtype=np.float32)
    rows, out = O.dedup_gradients(ids, g)
    assert rows.tolist() == [2, 5, 9]
    # (1e8 + -1e8) + 1 == 1 ; any other order of float adds gives 0
    assert out[:, 0].tolist() == [3.0, 1.0, 3.0]


def test_pcg32_known_answer_vector():
    """pcg32-demo (pcg-random.org, pcg32_srandom_r(42, 54)) round-1 outputs; the sampler masks the sign bit."""
    kat = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    got = O.random_positive_ints(42, 54, 6)
    assert got.tolist() == [v & 0x7FFFFFFF for v in kat]


def _ref_cpu_sampler(r, M, N):
    # graph_sampling_test_utils.cu:306-321 with Offset = 0
    Q = list(range(N))
    a = []
    for i in range(M):
        a.append(Q[r[i]])
        Q[r[i]] = Q[N - i - 1]
    return a


def test_selection_matches_reference_cpu_sampler():
    rng = np.random.default_rng(9)
    for _ in range(200):
        N = int(rng.integers(2, 300))
        M = int(rng.integers(1, N))
        r = np.array([rng.integers(0, N - i) for i in range(M)], dtype=np.int32)
        got = O.fisher_yates(r, M, N).tolist()
        assert got == _ref_cpu_sampler(r.tolist(), M, N)
        assert len(set(got)) == M and all(0 <= x < N for x in got)


def test_sampler_shape_table():
    # unweighted_sample_without_replacement_func.cuh:423-458
    assert O.sampler_shape(10) == (32, 1) and O.sampler_shape(25) == (32, 1) and O.sampler_shape(32) == (32, 1)
    assert O.sampler_shape(33) == (32, 2) and O.sampler_shape(96) == (32, 3) and O.sampler_shape(97) == (64, 2)
    assert O.sampler_shape(193) == (128, 2) and O.sampler_shape(385) == (256, 2) and O.sampler_shape(513) == (256, 3)
    assert O.sampler_shape(769) == (256, 4) and O.sampler_shape(1024) == (256, 4)


def test_unweighted_sample_structure():
    rng = np.random.default_rng(10)
    nodes = 300
    deg = rng.integers(0, 60, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, nodes, size=int(row_ptr[-1])).astype(np.int64)
    centers = rng.integers(0, nodes, size=128).astype(np.int64)
    for k in (10, 25, 40):
        off, dst, lid, gid = O.unweighted_sample(row_ptr, col, centers, k, 1234)
        cnt = np.minimum(deg[centers], k)
        assert np.array_equal(np.diff(off), cnt) and off[0] == 0
        for c, node in enumerate(centers):
            s, e = off[c], off[c + 1]
            lo, hi = row_ptr[node], row_ptr[node + 1]
            assert np.all(lid[s:e] == c) and np.all((gid[s:e] >= lo) & (gid[s:e] < hi))
            assert len(set(gid[s:e].tolist())) == e - s  # without replacement
            assert np.array_equal(dst[s:e], col[gid[s:e]])
            if deg[node] <= k:
                assert np.array_equal(gid[s:e], np.arange(lo, hi))  # CSR order when everything is taken
    off, dst, lid, gid = O.unweighted_sample(row_ptr, col, centers, -1, 1)
    assert np.array_equal(np.diff(off), deg[centers])


def test_weighted_sample_oracle_structure_and_bias():
    """A-Res restatement: samples are distinct neighbours; heavy edges are kept far more often than light ones."""
    rng = np.random.default_rng(21)
    nodes = 200
    deg = rng.integers(0, 90, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, nodes, size=int(row_ptr[-1])).astype(np.int64)
    w = rng.uniform(0.1, 1.0, size=col.size).astype(np.float32)
    heavy = rng.random(col.size) < 0.1
    w[heavy] = 50.0
    centers = rng.integers(0, nodes, size=300).astype(np.int64)
    kept_heavy = kept_light = tot_heavy = tot_light = 0
    for k, seed in ((10, 1), (25, 2), (40, 3)):
        off, dst, lid, gid, margin = O.weighted_sample(row_ptr, col, w, centers, k, seed)
        assert np.array_equal(np.diff(off), np.minimum(deg[centers], k))
        for c, node in enumerate(centers):
            s, e = off[c], off[c + 1]
            lo, hi = row_ptr[node], row_ptr[node + 1]
            assert np.all(lid[s:e] == c) and np.all((gid[s:e] >= lo) & (gid[s:e] < hi)) and len(set(gid[s:e].tolist())) == e - s
            assert np.array_equal(dst[s:e], col[gid[s:e]])
            if deg[node] <= k:
                assert np.array_equal(gid[s:e], np.arange(lo, hi)) and np.isinf(margin[c])
            else:
                sel = np.zeros(hi - lo, bool)
                sel[gid[s:e] - lo] = True
                hv = heavy[lo:hi]
                kept_heavy += int((sel & hv).sum()); tot_heavy += int(hv.sum())
                kept_light += int((sel & ~hv).sum()); tot_light += int((~hv).sum())
    assert kept_heavy / max(tot_heavy, 1) > 2.0 * kept_light / max(tot_light, 1)
    # the key stream: log2(u)/w, negative, and the weight-1 helper equals the k-stream of a unit-weight edge
    keys = O.exponential_negative_floats(5, 0, 1000)
    assert np.all(keys < 0) and np.all(np.isfinite(keys))
    assert abs(float(np.mean(keys)) + 1.0 / np.log(2.0)) < 0.15  # E[log2 U] = -1/ln 2
