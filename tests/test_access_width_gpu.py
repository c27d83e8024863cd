"""Every access width of the row-move kernel (32, 16, 8, 4, 2, 1 bytes) against the oracle, gather and scatter.

The kernel moves rows in the widest power-of-two unit that divides the row size, both row strides and every base address
(csrc/gather_scatter.cu: 256-bit LDG/STG where all of them are multiples of 32 bytes).  Each case below pins one width by
construction -- row bytes, table stride, output stride -- and large enough batches take the ~4 KiB-per-warp batching as
well as the small-call path.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

# (dtype, cols, table stride, out stride) -> widest unit in bytes
CASES = [
    (O.DT_FLOAT, 64, 64, 64, 32),     # 256 B rows, everything 32-aligned
    (O.DT_FLOAT, 256, 256, 256, 32),  # 1 KiB rows
    (O.DT_HALF, 128, 128, 128, 32),   # 256 B fp16 rows (config C3's row)
    (O.DT_FLOAT, 64, 68, 64, 16),     # table stride 272 B: 16-aligned only
    (O.DT_FLOAT, 60, 60, 60, 16),     # 240 B rows
    (O.DT_FLOAT, 62, 62, 62, 8),
    (O.DT_FLOAT, 63, 63, 63, 4),
    (O.DT_HALF, 63, 63, 63, 2),
    (O.DT_INT8, 63, 63, 63, 1),
    (O.DT_FLOAT, 64, 64, 65, 4),      # only the output stride is odd
]


@pytest.mark.parametrize("dt,cols,stride,out_stride,unit", CASES)
@pytest.mark.parametrize("n", [37, 70000])
@pytest.mark.parametrize("mem_type", ["continuous", "chunked"])
def test_gather_and_scatter_at_every_access_width(dt, cols, stride, out_stride, unit, n, mem_type):
    import gpu_utils as G
    esz = np.dtype(O.NP_OF[dt]).itemsize
    assert np.gcd.reduce([cols * esz, stride * esz, out_stride * esz, 32]) == unit  # the case pins the width it says it does
    rows = 100000
    rng = np.random.default_rng(cols * 7 + n + unit)
    comm = G.single_comm()
    table, view = G.create_table(comm, mem_type, "cuda", dt, rows, cols, stride)
    host = G.random_table(rng, dt, rows, stride)
    view.copy_(G.np_to_torch(host, dt))
    idx = rng.integers(0, rows, size=n).astype(np.int64)
    idx[::13] = -1
    sentinel = G.random_table(rng, dt, n, out_stride)
    out_t = G.np_to_torch(sentinel.copy(), dt).cuda()
    G.gather(table, G.idx_to_cuda(idx), out_t[:, :cols] if out_stride != cols else out_t)
    torch.cuda.synchronize()
    exp = sentinel.copy()
    O.gather(host, dt, idx, dt, out=exp, cols=cols)
    assert G.torch_to_np(out_t, dt).tobytes() == exp.tobytes(), "gather"
    # scatter distinct rows back from a strided source
    sidx = rng.permutation(rows)[:n].astype(np.int64)
    sidx[::17] = -1
    src = G.random_table(rng, dt, n, out_stride)
    src_t = G.np_to_torch(src, dt).cuda()
    G.scatter(src_t[:, :cols] if out_stride != cols else src_t, G.idx_to_cuda(sidx), table)
    torch.cuda.synchronize()
    O.scatter(src, dt, sidx, host, dt, cols=cols)
    got = G.torch_to_np(view, dt)
    G.wmb.destroy_wholememory_tensor(table)
    assert got[:, :cols].tobytes() == host[:, :cols].tobytes(), "scatter"


@pytest.mark.parametrize("dt,cols", [(O.DT_INT8, 100001), (O.DT_FLOAT, 300000), (O.DT_HALF, 70001)])
def test_very_long_rows(dt, cols):
    """Rows far beyond the batched-row map's range (round 1 refused more than 32,767 units per row): one row per warp,
    any width -- 100,001-byte int8 rows move in 1-byte units, 1.2 MB fp32 rows in 32-byte units."""
    import gpu_utils as G
    rows, n = 37, 50
    rng = np.random.default_rng(cols)
    comm = G.single_comm()
    table, view = G.create_table(comm, "chunked", "cuda", dt, rows, cols, cols)
    host = G.random_table(rng, dt, rows, cols)
    view.copy_(G.np_to_torch(host, dt))
    idx = rng.integers(0, rows, size=n).astype(np.int64)
    idx[5] = -1
    sentinel = G.random_table(rng, dt, n, cols)
    out_t = G.np_to_torch(sentinel.copy(), dt).cuda()
    G.gather(table, G.idx_to_cuda(idx), out_t)
    torch.cuda.synchronize()
    exp = sentinel.copy()
    O.gather(host, dt, idx, dt, out=exp)
    assert G.torch_to_np(out_t, dt).tobytes() == exp.tobytes(), "gather"
    sidx = rng.permutation(rows)[:20].astype(np.int64)
    src = G.random_table(rng, dt, 20, cols)
    G.scatter(G.np_to_torch(src, dt).cuda(), G.idx_to_cuda(sidx), table)
    torch.cuda.synchronize()
    O.scatter(src, dt, sidx, host, dt)
    got = G.torch_to_np(view, dt)
    G.wmb.destroy_wholememory_tensor(table)
    assert got.tobytes() == host.tobytes(), "scatter"
