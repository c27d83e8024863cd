"""BASELINE.json configs at their FULL single-GPU sizes, checked through size-independent properties.

The oracle cannot re-walk a 100 GB table in seconds, so the table holds the reference tests' closed-form pattern
(element (r, c) = r & mask, cpp/tests/wholememory_ops/embedding_test_utils.cu:197-238): every gathered row is
checkable from its index alone, a scatter followed by a gather must return the scattered bytes, and gathering twice must
give the same bytes.

  C1  wholememory_gather 1M x 64 fp32, single rank, HOST memory, 100,000 indices  == numpy table[idx]
  C2  1-GPU CONTINUOUS 100M x 256 fp32 (102.4 GB), 1,048,576 uniform int64 indices
  C3  one rank's shard of the 8-GPU CHUNKED 1B x 128 fp16 table: 125M x 128 fp16 (32 GB), same batch
  NS  one rank's shard of the north-star 1B x 256 fp16 table: 125M x 256 fp16 (64 GB)

(File name sorts last on purpose: these allocate most of the GPU and were added without a GPU at hand; the same
configurations are what bench.py times and asserts.)"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def G():
    import gpu_utils as g
    g.single_comm()
    return g


def _free_bytes():
    torch.cuda.empty_cache()
    return torch.cuda.mem_get_info()[0]


def _fill_pattern(view, mask, th_dtype, chunk=4_000_000):
    rows = view.shape[0]
    for s in range(0, rows, chunk):
        e = min(rows, s + chunk)
        ids = torch.arange(s, e, device=view.device, dtype=torch.int64)
        view[s:e] = (ids & mask).to(th_dtype).unsqueeze(1)


def test_c1_host_memory_gather_equals_numpy_take(G):
    rows, cols, n = 1_000_000, 64, 100_000
    comm = G.single_comm()
    table, view = G.create_table(comm, "continuous", "cpu", O.DT_FLOAT, rows, cols, cols)
    rng = np.random.default_rng(0x5EED)
    host = rng.standard_normal((rows, cols)).astype(np.float32)
    view.copy_(torch.from_numpy(host))
    for idt in (np.int64, np.int32):
        idx = rng.integers(0, rows, size=n).astype(idt)
        out = torch.zeros(n, cols, device="cuda")
        G.gather(table, torch.from_numpy(idx).cuda(), out)
        torch.cuda.synchronize()
        assert out.cpu().numpy().tobytes() == host[idx].tobytes()
    G.wmb.destroy_wholememory_tensor(table)


@pytest.mark.parametrize("name,mem_type,rows,cols,dt", [
    ("C2", "continuous", 100_000_000, 256, O.DT_FLOAT),
    ("C3-shard", "chunked", 125_000_000, 128, O.DT_HALF),
    ("NS-shard", "chunked", 125_000_000, 256, O.DT_HALF),
])
def test_full_size_gather_scatter_properties(G, name, mem_type, rows, cols, dt):
    esize = 4 if dt == O.DT_FLOAT else 2
    th = torch.float32 if dt == O.DT_FLOAT else torch.float16
    mask = (1 << 24) - 1 if dt == O.DT_FLOAT else (1 << 11) - 1  # integers every value of the dtype represents exactly
    need = rows * cols * esize + (6 << 30)
    if _free_bytes() < need:
        pytest.skip("%s needs %.0f GB of free HBM" % (name, need / 1e9))
    n = 1 << 20
    comm = G.single_comm()
    # same creation + local view calls as bench.py
    table = G.wmb.create_wholememory_matrix(G.WM_OF[dt], rows, cols, -1, comm, G.MT[mem_type], G.wmb.MlDevice)
    view, first_row = table.get_local_tensor(G.wmb.MlDevice, torch.cuda.current_device())
    assert first_row == 0 and tuple(view.shape) == (rows, cols)
    try:
        _fill_pattern(view, mask, th)
        g = torch.Generator(device="cuda")
        g.manual_seed(0x5EED)
        idx = torch.randint(0, rows, (n,), device="cuda", dtype=torch.int64, generator=g)
        idx[0], idx[1] = 0, rows - 1  # first and last row of the table
        out = torch.empty(n, cols, device="cuda", dtype=th)
        G.gather(table, idx, out)
        torch.cuda.synchronize()
        exp_col = (idx & mask).to(th)
        # every element of a gathered row equals f(row id): check the first, the last and the row sum
        assert torch.equal(out[:, 0], exp_col) and torch.equal(out[:, cols - 1], exp_col)
        assert torch.equal(out.sum(dim=1, dtype=torch.float64), exp_col.double() * cols)
        # idempotence
        out2 = torch.empty_like(out)
        G.gather(table, idx, out2)
        torch.cuda.synchronize()
        assert torch.equal(out, out2)
        # int32 indices reach the same rows (all ids < 2^31 here)
        out3 = torch.empty_like(out)
        G.gather(table, idx.to(torch.int32), out3)
        torch.cuda.synchronize()
        assert torch.equal(out, out3)
        # scatter -> gather round trip on distinct rows
        m = 1 << 16
        sidx = torch.unique(idx)[:m].contiguous()
        src = torch.randn(sidx.shape[0], cols, device="cuda", generator=g).to(th)
        G.scatter(src, sidx, table)
        back = torch.empty_like(src)
        G.gather(table, sidx, back)
        torch.cuda.synchronize()
        assert torch.equal(back.view(torch.int16 if esize == 2 else torch.int32), src.view(torch.int16 if esize == 2 else torch.int32))
        # neighbours of the scattered rows are untouched
        nb = torch.clamp(sidx + 1, max=rows - 1)
        keep = ~torch.isin(nb, sidx)
        nb = nb[keep].contiguous()
        chk = torch.empty(nb.shape[0], cols, device="cuda", dtype=th)
        G.gather(table, nb, chk)
        torch.cuda.synchronize()
        assert torch.equal(chk[:, 0], (nb & mask).to(th)) and torch.equal(chk[:, cols - 1], (nb & mask).to(th))
    finally:
        del view
        G.wmb.destroy_wholememory_tensor(table)
        torch.cuda.empty_cache()
