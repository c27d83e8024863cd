"""A 1-D WholeMemory array with the dense operand the REFERENCE expects for it: the 2-D [n, 1] matrix (gather_op.cpp:44-58 and
scatter_op.cpp:44-58 unsqueeze the table to [N, 1] before comparing ranks).  The 1-D dense form, which this library also
accepts, is covered by the verified test_gather_scatter_gpu.py; the [n, 1] form was aligned with the reference on CPU
(tests/test_ref_host_ops.py) after the GPU budget was spent, so this is its first execution on the device.

(File name sorts last on purpose.)"""
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("mem_type", ["continuous", "chunked", "distributed"])
@pytest.mark.parametrize("wm_dt,th_dt", [("DtInt64", torch.int64), ("DtFloat", torch.float32), ("DtInt", torch.int32)])
def test_one_dim_table_with_column_matrix_operands(mem_type, wm_dt, th_dt):
    import gpu_utils as G
    wmb = G.wmb
    comm = G.single_comm()
    n = 5000
    arr = wmb.create_wholememory_array(getattr(wmb, wm_dt), n, comm, G.MT[mem_type], wmb.MlDevice)
    try:
        flat, _ = arr.get_wholememory_handle().get_local_flatten_tensor(getattr(wmb, wm_dt), wmb.MlDevice, torch.cuda.current_device())
        vals = (torch.arange(n, device="cuda") * 3 + 1).to(th_dt)
        flat.copy_(vals)
        idx = torch.tensor([5, n - 1, 0, 17, -1, 4242], dtype=torch.int64, device="cuda")
        out = torch.full((idx.shape[0], 1), 7, dtype=th_dt, device="cuda")
        G.gather(arr, idx, out)
        torch.cuda.synchronize()
        exp = torch.where(idx >= 0, vals[idx.clamp(min=0)], torch.tensor(7, device="cuda").to(th_dt)).reshape(-1, 1)
        assert torch.equal(out, exp)                      # the negative index leaves its row untouched
        sidx = torch.tensor([9, 1, 4000], dtype=torch.int32, device="cuda")
        src = torch.tensor([[11], [22], [33]], device="cuda").to(th_dt)
        G.scatter(src, sidx, arr)
        torch.cuda.synchronize()
        assert flat[sidx.long()].tolist() == src.flatten().tolist() and flat[10].item() == vals[10].item()
        with pytest.raises((ValueError, RuntimeError)):   # a wider dense operand does not match the single column
            G.gather(arr, idx, torch.zeros(idx.shape[0], 2, dtype=th_dt, device="cuda"))
    finally:
        wmb.destroy_wholememory_tensor(arr)
