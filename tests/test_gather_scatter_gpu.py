"""GPU parity tests of wholememory_gather / wholememory_scatter against the oracle, through the C ABI.

Grid modelled on the reference's gather/scatter suites
(cpp/tests/wholememory_ops/wholememory_gather_tests.cu:288-527, wholememory_scatter_tests.cu:305-...):
memory types x locations x embedding dims {1,11,32,127,128,129,513} x strides x dtype pairs x index
dtypes x counts {0, N, N+5}, plus negative indices, column windows and strided outputs.
Bit-exact (raw bytes) in every case.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

DIMS = [1, 11, 32, 127, 128, 129, 513]


@pytest.fixture(scope="module")
def env():
    import gpu_utils as G
    return G


def _check_gather(G, mem_type, location, tab_dt, out_dt, idx_np, rows, cols, stride, out_stride=None, pattern=False,
                  seed=0, sms=-1):
    rng = np.random.default_rng(seed)
    comm = G.single_comm()
    table, view = G.create_table(comm, mem_type, location, tab_dt, rows, cols, stride)
    host = O.test_pattern(tab_dt, 0, rows, stride, stride) if pattern else G.random_table(rng, tab_dt, rows, stride)
    view.copy_(G.np_to_torch(host, tab_dt))
    torch.cuda.synchronize()
    n = idx_np.shape[0]
    out_stride = cols if out_stride is None else out_stride
    sentinel = G.random_table(rng, out_dt, max(n, 1), out_stride)[:n]
    out_t = G.np_to_torch(sentinel.copy(), out_dt).cuda()
    idx_t = G.idx_to_cuda(idx_np)
    G.gather(table, idx_t, out_t[:, :cols] if out_stride != cols else out_t, sms)
    torch.cuda.synchronize()
    got = G.torch_to_np(out_t, out_dt)
    exp = sentinel.copy()
    if n:
        O.gather(host, tab_dt, idx_np, out_dt, out=exp, cols=cols)
    G.wmb.destroy_wholememory_tensor(table)
    assert got.tobytes() == exp.tobytes(), f"mismatch rows={np.nonzero((got != exp).any(axis=1))[0][:8]}"


@pytest.mark.parametrize("cols", DIMS)
@pytest.mark.parametrize("tab_dt,out_dt", [(O.DT_FLOAT, O.DT_FLOAT), (O.DT_HALF, O.DT_HALF), (O.DT_FLOAT, O.DT_HALF),
                                           (O.DT_HALF, O.DT_FLOAT)])
@pytest.mark.parametrize("idx_dtype", [np.int32, np.int64])
def test_gather_continuous_device_dims(env, cols, tab_dt, out_dt, idx_dtype):
    rows = 3001
    rng = np.random.default_rng(cols)
    idx = rng.integers(0, rows, size=2005).astype(idx_dtype)
    _check_gather(env, "continuous", "cuda", tab_dt, out_dt, idx, rows, cols, cols, seed=cols)


@pytest.mark.parametrize("mem_type,location", [("continuous", "cuda"), ("chunked", "cuda"), ("distributed", "cuda"),
                                               ("continuous", "cpu"), ("chunked", "cpu"), ("distributed", "cpu")])
@pytest.mark.parametrize("cols,stride", [(11, 12), (32, 33), (128, 128), (129, 136)])
@pytest.mark.parametrize("count", [0, 1000, 1005])
def test_gather_memory_types_strides_counts(env, mem_type, location, cols, stride, count):
    rows = 2048
    rng = np.random.default_rng(count + cols)
    idx = rng.integers(0, rows, size=count).astype(np.int64)
    _check_gather(env, mem_type, location, O.DT_FLOAT, O.DT_FLOAT, idx, rows, cols, stride, pattern=True)


@pytest.mark.parametrize("tab_dt,out_dt", [p for p in itertools.product(O.FLOAT_DTS, O.FLOAT_DTS)] +
                         [p for p in itertools.product(O.INT_DTS, O.INT_DTS)])
def test_gather_every_dtype_pair(env, tab_dt, out_dt):
    rows, cols = 777, 24
    rng = np.random.default_rng(tab_dt * 10 + out_dt)
    idx = rng.integers(0, rows, size=500).astype(np.int64)
    _check_gather(env, "chunked", "cuda", tab_dt, out_dt, idx, rows, cols, cols, seed=tab_dt * 10 + out_dt)


def test_gather_reference_closed_form_pattern_fp16_and_bf16(env):
    for dt in (O.DT_HALF, O.DT_BF16, O.DT_DOUBLE, O.DT_INT8):
        idx = np.random.default_rng(1).integers(0, 5000, size=3000).astype(np.int64)
        _check_gather(env, "continuous", "cuda", dt, dt, idx, 5000, 33, 40, pattern=True)


def test_gather_negative_indices_leave_rows_untouched(env):
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 1000, size=777).astype(np.int64)
    idx[rng.integers(0, 777, size=100)] = -1
    idx[0] = -5
    _check_gather(env, "continuous", "cuda", O.DT_FLOAT, O.DT_FLOAT, idx, 1000, 64, 64)
    _check_gather(env, "chunked", "cuda", O.DT_HALF, O.DT_FLOAT, idx.astype(np.int32), 1000, 30, 32)


def test_gather_strided_output_and_sm_budget(env):
    idx = np.random.default_rng(3).integers(0, 4000, size=5000).astype(np.int64)
    _check_gather(env, "continuous", "cuda", O.DT_FLOAT, O.DT_FLOAT, idx, 4000, 256, 256, out_stride=300)
    _check_gather(env, "continuous", "cuda", O.DT_FLOAT, O.DT_FLOAT, idx, 4000, 256, 256, sms=7)
    _check_gather(env, "continuous", "cuda", O.DT_HALF, O.DT_HALF, idx, 4000, 2048, 2048)  # long rows: small batches


def test_gather_column_window_and_1d(env):
    G = env
    comm = G.single_comm()
    rows, stride = 500, 48
    rng = np.random.default_rng(4)
    table, view = G.create_table(comm, "chunked", "cuda", O.DT_FLOAT, rows, stride, stride)
    host = G.random_table(rng, O.DT_FLOAT, rows, stride)
    view.copy_(torch.from_numpy(host))
    sub = table.get_sub_tensor([0, 5], [-1, 30])  # columns [5, 30) of every row: odd offset => scalar path
    idx = rng.integers(0, rows, size=300).astype(np.int64)
    out = torch.zeros(300, 25, device="cuda")
    G.gather(sub, torch.from_numpy(idx).cuda(), out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host[idx, 5:30])
    G.wmb.destroy_wholememory_tensor(sub)
    G.wmb.destroy_wholememory_tensor(table)
    # 1-D WholeMemory array gathered into a 1-D output (gather_op.cpp:44-49 unsqueezes both)
    arr = G.wmb.create_wholememory_array(G.wmb.DtInt64, 1000, comm, G.wmb.MtContinuous, G.wmb.MlDevice)
    flat, _ = arr.get_wholememory_handle().get_local_flatten_tensor(G.wmb.DtInt64, G.wmb.MlDevice, 0)
    vals = torch.arange(1000, dtype=torch.int64, device="cuda") * 3 + 1
    flat.copy_(vals)
    idx_t = torch.tensor([5, 999, 0, 17], dtype=torch.int32, device="cuda")
    out1 = torch.zeros(4, dtype=torch.int64, device="cuda")
    G.gather(arr, idx_t, out1)
    torch.cuda.synchronize()
    assert out1.tolist() == [16, 2998, 1, 52]
    G.wmb.destroy_wholememory_tensor(arr)


def test_gather_argument_errors(env):
    G = env
    comm = G.single_comm()
    table, _ = G.create_table(comm, "continuous", "cuda", O.DT_FLOAT, 100, 8, 8)
    idx = torch.zeros(10, dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError):  # float table -> int output (gather_func.cu:78-81)
        G.gather(table, idx, torch.zeros(10, 8, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError):  # output rows != index count
        G.gather(table, idx, torch.zeros(9, 8, device="cuda"))
    with pytest.raises(ValueError):  # 2-D indices
        G.gather(table, torch.zeros(5, 2, dtype=torch.int64, device="cuda"), torch.zeros(10, 8, device="cuda"))
    with pytest.raises(ValueError):  # output rank differs from table rank
        G.gather(table, idx, torch.zeros(10, device="cuda"))
    G.wmb.destroy_wholememory_tensor(table)


@pytest.mark.parametrize("mem_type,location", [("continuous", "cuda"), ("chunked", "cuda"), ("distributed", "cuda"),
                                               ("chunked", "cpu")])
@pytest.mark.parametrize("cols,stride", [(1, 1), (11, 12), (128, 128), (513, 520)])
@pytest.mark.parametrize("in_dt,tab_dt", [(O.DT_FLOAT, O.DT_FLOAT), (O.DT_FLOAT, O.DT_HALF), (O.DT_INT64, O.DT_INT)])
def test_scatter_then_gather(env, mem_type, location, cols, stride, in_dt, tab_dt):
    """Reference scatter test: scatter rows, read the table back, compare raw bits (scatter_tests.cu:238-281)."""
    G = env
    comm = G.single_comm()
    rows, n = 1500, 1000
    rng = np.random.default_rng(cols * 7 + in_dt)
    table, view = G.create_table(comm, mem_type, location, tab_dt, rows, cols, stride)
    base = G.random_table(rng, tab_dt, rows, stride)
    view.copy_(G.np_to_torch(base, tab_dt))
    idx = rng.permutation(rows)[:n].astype(np.int64)  # unique: duplicates race in the reference too
    idx[::97] = -1
    src = G.random_table(rng, in_dt, n, cols)
    G.scatter(G.np_to_torch(src, in_dt).cuda(), torch.from_numpy(idx).cuda(), table)
    torch.cuda.synchronize()
    exp = base.copy()
    O.scatter(src, in_dt, idx, exp, tab_dt, cols=cols)
    got = G.torch_to_np(view, tab_dt)
    G.wmb.destroy_wholememory_tensor(table)
    assert got.tobytes() == exp.tobytes()


def test_scatter_duplicates_with_identical_rows(env):
    G = env
    comm = G.single_comm()
    table, view = G.create_table(comm, "continuous", "cuda", O.DT_FLOAT, 64, 16, 16)
    view.zero_()
    idx = torch.tensor([3, 3, 3, 9, 9], dtype=torch.int64, device="cuda")
    src = torch.arange(16, dtype=torch.float32, device="cuda").repeat(5, 1)
    G.scatter(src, idx, table)
    torch.cuda.synchronize()
    assert torch.equal(view[3], src[0]) and torch.equal(view[9], src[0]) and view[4].abs().sum() == 0
    G.wmb.destroy_wholememory_tensor(table)


def test_gather_large_property_checks(env):
    """BASELINE-sized rows (fp32 x 256) on a table too big for the oracle to re-walk cheaply: the closed-form
    pattern makes every gathered row checkable from its index alone (size-independent property)."""
    G = env
    comm = G.single_comm()
    rows, cols = 2_000_000, 256
    table, view = G.create_table(comm, "continuous", "cuda", O.DT_FLOAT, rows, cols, cols)
    ids = torch.arange(rows, device="cuda", dtype=torch.int64)
    view.copy_((ids & ((1 << 24) - 1)).to(torch.float32).unsqueeze(1).expand(rows, cols))
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED)
    idx = torch.randint(0, rows, (1 << 20,), device="cuda", generator=g)
    out = torch.empty(1 << 20, cols, device="cuda")
    G.gather(table, idx, out)
    torch.cuda.synchronize()
    exp_col = (idx & ((1 << 24) - 1)).to(torch.float32)
    assert torch.equal(out[:, 0], exp_col) and torch.equal(out[:, cols - 1], exp_col)
    assert torch.equal(out.sum(dim=1, dtype=torch.float64), exp_col.double() * cols)
    # idempotence: gathering twice gives the same bytes
    out2 = torch.empty_like(out)
    G.gather(table, idx, out2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    G.wmb.destroy_wholememory_tensor(table)


def test_scatter_and_gather_replay_from_a_cuda_graph(env):
    """The mapped-memory ops allocate nothing and never touch the host after launch, so a training step can capture them
    in a CUDA graph and replay it with new indices / rows in the same buffers (B200 launch-bound small-batch regime)."""
    G = env
    comm = G.single_comm()
    rows, cols, n = 4096, 96, 1500
    rng = np.random.default_rng(77)
    table, view = G.create_table(comm, "chunked", "cuda", O.DT_FLOAT, rows, cols, cols)
    base = G.random_table(rng, O.DT_FLOAT, rows, cols)
    view.copy_(G.np_to_torch(base, O.DT_FLOAT))
    src_t = torch.zeros(n, cols, device="cuda")
    sidx_t = torch.zeros(n, dtype=torch.int64, device="cuda")
    gidx_t = torch.zeros(n, dtype=torch.int32, device="cuda")
    out_t = torch.zeros(n, cols, dtype=torch.float16, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):  # warm-up outside capture (one-time lazy initialisation)
        G.scatter(src_t, sidx_t, table)  # all-zero rows onto row 0; the table is restored below
        G.gather(table, gidx_t, out_t)
    side.synchronize()
    view.copy_(G.np_to_torch(base, O.DT_FLOAT))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        G.scatter(src_t, sidx_t, table)
        G.gather(table, gidx_t, out_t)
    exp_table = base.copy()
    for trial in range(3):
        sidx = rng.permutation(rows)[:n].astype(np.int64)
        gidx = rng.integers(0, rows, size=n).astype(np.int32)
        src = G.random_table(rng, O.DT_FLOAT, n, cols)
        src_t.copy_(torch.from_numpy(src))
        sidx_t.copy_(torch.from_numpy(sidx))
        gidx_t.copy_(torch.from_numpy(gidx))
        graph.replay()
        torch.cuda.synchronize()
        O.scatter(src, O.DT_FLOAT, sidx, exp_table, O.DT_FLOAT, cols=cols)
        exp_out = np.zeros((n, cols), np.float16)
        O.gather(exp_table, O.DT_FLOAT, gidx, O.DT_HALF, out=exp_out, cols=cols)
        assert G.torch_to_np(view, O.DT_FLOAT).tobytes() == exp_table.tobytes(), f"table differs after replay {trial}"
        assert G.torch_to_np(out_t, O.DT_HALF).tobytes() == exp_out.tobytes(), f"gathered rows differ after replay {trial}"
    del graph
    G.wmb.destroy_wholememory_tensor(table)
