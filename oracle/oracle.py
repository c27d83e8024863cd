"""ORACLE -- TEST INFRASTRUCTURE ONLY (see the header of wm_oracle.c).

ctypes front-end to ``oracle/_build/liboracle.so`` (the plain-C restatement of the reference's hot
path) plus a second, independent numpy restatement of gather used to cross-check the C one.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

# numeric values of wholememory_dtype_t (reference tensor_description.h:29-40)
DT_FLOAT, DT_HALF, DT_DOUBLE, DT_BF16, DT_INT, DT_INT64, DT_INT16, DT_INT8 = 1, 2, 3, 4, 5, 6, 7, 8

# bf16 has no numpy dtype: it travels as uint16 bit patterns
NP_OF = {DT_FLOAT: np.float32, DT_HALF: np.float16, DT_DOUBLE: np.float64, DT_BF16: np.uint16,
         DT_INT: np.int32, DT_INT64: np.int64, DT_INT16: np.int16, DT_INT8: np.int8}
FLOAT_DTS = (DT_FLOAT, DT_HALF, DT_DOUBLE, DT_BF16)
INT_DTS = (DT_INT, DT_INT64, DT_INT16, DT_INT8)


def build():
    """Compile the C oracle (gcc, a second or two)."""
    srcs = [os.path.join(_HERE, f) for f in ("wm_oracle.c", "wm_oracle_weighted.c")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        i64, vp, ci, f32 = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        _lib.oracle_gather.argtypes = [vp, ci, i64, i64, i64, vp, ci, i64, vp, ci, i64, i64]
        _lib.oracle_scatter.argtypes = [vp, ci, i64, i64, i64, vp, ci, i64, vp, ci, i64, i64]
        _lib.oracle_convert.argtypes = [vp, ci, vp, ci, i64]
        _lib.oracle_fill_test_pattern.argtypes = [vp, ci, i64, i64, i64, i64]
        _lib.oracle_partition.argtypes = [i64, ci, vp]
        _lib.oracle_sgd.argtypes = [vp, i64, i64, vp, i64, vp, i64, f32, f32]
        _lib.oracle_lazy_adam.argtypes = [vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, i64, f32, f32, f32, f32, ci, f32]
        _lib.oracle_adagrad.argtypes = [vp, i64, vp, i64, i64, vp, i64, vp, i64, f32, f32, f32]
        _lib.oracle_rmsprop.argtypes = [vp, i64, vp, i64, i64, vp, i64, vp, i64, f32, f32, f32, f32]
        _lib.oracle_dedup_gradients.argtypes = [vp, i64, vp, i64, i64, vp, vp]
        _lib.oracle_dedup_gradients.restype = i64
        _lib.oracle_random_positive_ints.argtypes = [ctypes.c_uint64, ctypes.c_uint64, vp, i64]
        _lib.oracle_fisher_yates.argtypes = [vp, ci, ci, vp]
        _lib.oracle_sampler_shape.argtypes = [ci, vp, vp]
        _lib.oracle_unweighted_sample.argtypes = [vp, vp, vp, i64, ci, ctypes.c_uint64, vp, vp, vp, vp]
        _lib.oracle_unweighted_sample.restype = i64
        _lib.oracle_gather_mt.argtypes = [vp, i64, i64, vp, i64, vp, i64, ci]
        _lib.oracle_exponential_negative_floats.argtypes = [ctypes.c_uint64, ctypes.c_uint64, vp, i64]
        _lib.oracle_weighted_sample.argtypes = [vp, vp, vp, ci, vp, i64, ci, ctypes.c_uint64, vp, vp, vp, vp, vp]
        _lib.oracle_weighted_sample.restype = i64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _idx_dt(idx):
    assert idx.dtype in (np.int32, np.int64)
    return DT_INT if idx.dtype == np.int32 else DT_INT64


def gather(table, tab_dt, idx, out_dt, out=None, cols=None, tab_off=0, out_stride=None, out_off=0):
    """out[i] = convert(table[idx[i]]); table is a 2-D C-contiguous numpy array whose row length is the stride."""
    table = np.ascontiguousarray(table)
    tab_stride = table.shape[1]
    cols = tab_stride - (tab_off % tab_stride) if cols is None else cols
    n = idx.shape[0]
    if out is None:
        out_stride = cols if out_stride is None else out_stride
        out = np.zeros((n, out_stride), dtype=NP_OF[out_dt])
    else:
        out_stride = out.shape[1]
    idx = np.ascontiguousarray(idx)
    lib().oracle_gather(_p(table), tab_dt, tab_stride, tab_off, cols, _p(idx), _idx_dt(idx), n, _p(out), out_dt,
                        out_stride, out_off)
    return out


def scatter(inp, in_dt, idx, table, tab_dt, cols=None, tab_off=0, in_off=0):
    """table[idx[i]] = convert(inp[i]) in place."""
    assert inp.flags["C_CONTIGUOUS"] and table.flags["C_CONTIGUOUS"]
    cols = inp.shape[1] if cols is None else cols
    idx = np.ascontiguousarray(idx)
    lib().oracle_scatter(_p(inp), in_dt, inp.shape[1], in_off, cols, _p(idx), _idx_dt(idx), idx.shape[0], _p(table),
                         tab_dt, table.shape[1], tab_off)
    return table


def convert(src, src_dt, dst_dt):
    src = np.ascontiguousarray(src)
    dst = np.zeros(src.shape, dtype=NP_OF[dst_dt])
    lib().oracle_convert(_p(src), src_dt, _p(dst), dst_dt, src.size)
    return dst


def test_pattern(dt, first_row, rows, cols, stride=None):
    """The reference tests' closed-form table (embedding_test_utils.cu:197-238)."""
    stride = cols if stride is None else stride
    t = np.zeros((rows, stride), dtype=NP_OF[dt])
    lib().oracle_fill_test_pattern(_p(t), dt, first_row, rows, cols, stride)
    return t


def partition(entries, world_size):
    off = np.zeros(world_size + 1, dtype=np.int64)
    lib().oracle_partition(entries, world_size, _p(off))
    return off


# ---------------------------------------------------------------- independent numpy restatement
def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _f32_to_bf16(f32):
    x = np.ascontiguousarray(f32, dtype=np.float32).view(np.uint32)
    nan = (x & 0x7FFFFFFF) > 0x7F800000
    r = ((x + 0x7FFF + ((x >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def np_convert(a, src_dt, dst_dt):
    """type_caster chain (gather_scatter_func.cuh:161-208): fp16/bf16 pass through float32."""
    if src_dt == dst_dt:
        return a.copy()
    if src_dt in INT_DTS:
        return a.astype(NP_OF[dst_dt])  # C truncation semantics == numpy astype for ints
    # load
    if src_dt == DT_BF16:
        v = _bf16_to_f32(a)
    elif src_dt == DT_HALF:
        v = a.astype(np.float32)
    else:
        v = a
    # store
    if dst_dt in (DT_HALF, DT_BF16):
        v32 = v.astype(np.float32)  # double -> float first (double rounding, as the reference)
        return v32.astype(np.float16) if dst_dt == DT_HALF else _f32_to_bf16(v32)
    return v.astype(NP_OF[dst_dt])


def np_gather(table, tab_dt, idx, out_dt, cols=None, col_off=0, out=None):
    """numpy-index restatement: rows with idx < 0 keep the previous content of `out` (zeros by default)."""
    cols = table.shape[1] - col_off if cols is None else cols
    n = idx.shape[0]
    res = np.zeros((n, cols), dtype=NP_OF[out_dt]) if out is None else out
    ok = idx >= 0
    res[ok, :cols] = np_convert(table[idx[ok], col_off:col_off + cols], tab_dt, out_dt)
    return res


# ---------------------------------------------------------------- optimizers
def optimizer_step(kind, w, rows, grads, lr, state=None, b12=None, weight_decay=0.0, epsilon=1e-8, beta1=0.9,
                   beta2=0.999, adam_w=False, alpha=0.99, dim=None):
    """One step on unique `rows` (int64) with fp32 grads [n, >=dim]; w/state are 2-D fp32 arrays, updated in place.
    state for adam = (m, v) arrays with the same row stride; adagrad/rmsprop = single array."""
    dim = grads.shape[1] if dim is None else dim
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    grads = np.ascontiguousarray(grads, dtype=np.float32)
    L = lib()
    n = rows.shape[0]
    if kind == "sgd":
        L.oracle_sgd(_p(w), w.shape[1], dim, _p(rows), n, _p(grads), grads.shape[1], weight_decay, lr)
    elif kind == "adam":
        m, v = state
        L.oracle_lazy_adam(_p(w), w.shape[1], _p(m), _p(v), m.shape[1], _p(b12), dim, _p(rows), n, _p(grads),
                           grads.shape[1], weight_decay, epsilon, beta1, beta2, int(adam_w), lr)
    elif kind == "adagrad":
        L.oracle_adagrad(_p(w), w.shape[1], _p(state), state.shape[1], dim, _p(rows), n, _p(grads), grads.shape[1],
                         weight_decay, epsilon, lr)
    elif kind == "rmsprop":
        L.oracle_rmsprop(_p(w), w.shape[1], _p(state), state.shape[1], dim, _p(rows), n, _p(grads), grads.shape[1],
                         weight_decay, epsilon, alpha, lr)
    else:
        raise ValueError(kind)


def dedup_gradients(ids, grads, dim=None):
    """(unique ids ascending, summed grads) -- duplicates added in arrival order (stable)."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    grads = np.ascontiguousarray(grads, dtype=np.float32)
    dim = grads.shape[1] if dim is None else dim
    out_rows = np.zeros(ids.shape[0], dtype=np.int64)
    out_g = np.zeros((ids.shape[0], dim), dtype=np.float32)
    u = lib().oracle_dedup_gradients(_p(ids), ids.shape[0], _p(grads), grads.shape[1], dim, _p(out_rows), _p(out_g))
    return out_rows[:u].copy(), out_g[:u].copy()


# ---------------------------------------------------------------- sampler
def random_positive_ints(seed, subsequence, count):
    out = np.zeros(count, dtype=np.int32)
    lib().oracle_random_positive_ints(seed, subsequence, _p(out), count)
    return out


def fisher_yates(r, M, N):
    r = np.ascontiguousarray(r, dtype=np.int32)
    out = np.zeros(M, dtype=np.int32)
    lib().oracle_fisher_yates(_p(r), M, N, _p(out))
    return out


def sampler_shape(k):
    b, i = ctypes.c_int(), ctypes.c_int()
    lib().oracle_sampler_shape(k, ctypes.byref(b), ctypes.byref(i))
    return b.value, i.value


def unweighted_sample(row_ptr, col, centers, k, seed):
    """Returns (offsets int32[n+1], dst int64[S], center_local int32[S], edge_gid int64[S])."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    centers = np.ascontiguousarray(centers, dtype=np.int64)
    n = centers.shape[0]
    offsets = np.zeros(n + 1, dtype=np.int32)
    total = lib().oracle_unweighted_sample(_p(row_ptr), _p(col), _p(centers), n, k, seed, _p(offsets), None, None, None)
    dst = np.zeros(total, dtype=np.int64)
    lid = np.zeros(total, dtype=np.int32)
    gid = np.zeros(total, dtype=np.int64)
    lib().oracle_unweighted_sample(_p(row_ptr), _p(col), _p(centers), n, k, seed, _p(offsets), _p(dst), _p(lid), _p(gid))
    return offsets, dst, lid, gid


def exponential_negative_floats(seed, subsequence, count):
    out = np.zeros(count, dtype=np.float32)
    lib().oracle_exponential_negative_floats(seed, subsequence, _p(out), count)
    return out


def weighted_sample(row_ptr, col, weights, centers, k, seed):
    """Returns (offsets, dst, center_local, edge_gid, margin) -- margin[c] = relative gap between the k-th and (k+1)-th
    best key of center c (inf when every neighbour is taken): a tiny margin marks a last-ulp tie."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    is_float = 1 if weights.dtype == np.float32 else 0
    w = np.ascontiguousarray(weights, dtype=np.float64)
    centers = np.ascontiguousarray(centers, dtype=np.int64)
    n = centers.shape[0]
    offsets = np.zeros(n + 1, dtype=np.int32)
    total = lib().oracle_weighted_sample(_p(row_ptr), _p(col), _p(w), is_float, _p(centers), n, k, seed, _p(offsets), None,
                                         None, None, None)
    dst = np.zeros(total, dtype=np.int64)
    lid = np.zeros(total, dtype=np.int32)
    gid = np.zeros(total, dtype=np.int64)
    margin = np.zeros(n, dtype=np.float32)
    lib().oracle_weighted_sample(_p(row_ptr), _p(col), _p(w), is_float, _p(centers), n, k, seed, _p(offsets), _p(dst), _p(lid),
                                 _p(gid), _p(margin))
    return offsets, dst, lid, gid, margin


# ---------------------------------------------------------------- timed CPU baseline
def gather_mt(table2d, idx, out2d, threads):
    """Multithreaded same-dtype row gather (bench.py cpu_baseline leg)."""
    assert table2d.flags["C_CONTIGUOUS"] and out2d.flags["C_CONTIGUOUS"] and idx.dtype == np.int64
    row_bytes = min(table2d.shape[1], out2d.shape[1]) * table2d.itemsize
    lib().oracle_gather_mt(_p(table2d), table2d.strides[0], row_bytes, _p(idx), idx.shape[0], _p(out2d),
                           out2d.strides[0], threads)
    return out2d
