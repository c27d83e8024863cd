"""Drop-in proof, second half (own process): the REFERENCE's unmodified cython module
(oracle/_ref/refbinding/wholememory_binding*.so, linked to THIS repo's libwholegraph.so) drives a LazyAdam training step,
unweighted neighbor sampling, append_unique and csr_add_self_loop; every result is checked against the oracle.
The env functions are a transcription of what pylibwholegraph/torch/wholegraph_env.py registers."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "refbinding"))


class _Ctx(object):
    def __init__(self):
        self.tensor = None


def main():
    import wholememory_binding as rwmb  # the reference's module
    from oracle import oracle as O

    torch.cuda.set_device(0)
    rwmb.init(0)
    comm = rwmb.create_communicator(rwmb.create_unique_id(), 0, 1)
    DT = rwmb.WholeMemoryDataType
    TH = {DT.DtFloat: torch.float32, DT.DtInt: torch.int32, DT.DtInt64: torch.int64, DT.DtInt8: torch.int8, DT.DtHalf: torch.float16}
    WM = {v: k for k, v in TH.items()}

    def create_ctx(global_context):
        return _Ctx()

    def destroy_ctx(memory_context, global_context):
        memory_context.tensor = None

    def malloc_fn(tensor_desc, malloc_type, memory_context, global_context):
        if malloc_type.get_type() == rwmb.WholeMemoryMemoryAllocType.MatDevice:
            t = torch.empty(tensor_desc.shape, dtype=TH[tensor_desc.dtype], device="cuda")
        else:
            t = torch.empty(tensor_desc.shape, dtype=TH[tensor_desc.dtype],
                            pin_memory=malloc_type.get_type() == rwmb.WholeMemoryMemoryAllocType.MatPinned)
        memory_context.tensor = t
        return t.data_ptr()

    def free_fn(memory_context, global_context):
        memory_context.tensor = None

    gctx = object()
    env = rwmb.GlobalContextWrapper()
    env.create_context(create_ctx, destroy_ctx, malloc_fn, free_fn, gctx, malloc_fn, free_fn, gctx)
    ctypes.pythonapi.Py_IncRef(ctypes.py_object(env))  # same lifetime as in the reference: never collected

    def wrap(t):
        d = rwmb.PyWholeMemoryTensorDescription()
        d.set_dtype(WM[t.dtype])
        d.set_storage_offset(0)
        d.set_shape(tuple(t.shape))
        d.set_stride(tuple(t.stride()))
        return rwmb.WrappedLocalTensor().wrap_tensor(d, t.data_ptr())

    def from_dlpack(dp):
        return torch.utils.dlpack.from_dlpack(dp.__dlpack__())

    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(21)

    # ---- embedding + LazyAdam: two steps with duplicate ids
    rows, dim, n = 2000, 100, 1500
    d = rwmb.PyWholeMemoryTensorDescription()
    d.set_dtype(DT.DtFloat)
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = rwmb.create_embedding(d, comm, rwmb.WholeMemoryMemoryType.MtChunked, rwmb.WholeMemoryMemoryLocation.MlDevice,
                                rwmb.create_non_cache_policy(), embedding_entry_partition=None, user_defined_sms=-1, round_robin_size=0)
    opt = rwmb.create_optimizer(rwmb.WholeMemoryOptimizerType.OptLazyAdam, {"beta1": 0.85, "weight_decay": 0.01})
    opt.add_embedding(emb)
    assert emb.get_optimizer_state_names() == ["m", "v", "beta12t"]
    wt = emb.get_embedding_tensor()
    local, off = wt.get_local_tensor(from_dlpack, rwmb.WholeMemoryMemoryLocation.MlDevice, 0)
    assert off == 0 and tuple(local.shape) == (rows, dim)
    w = rng.standard_normal((rows, dim)).astype(np.float32)
    local.copy_(torch.from_numpy(w))
    m, v, b12 = np.zeros_like(w), np.zeros_like(w), np.ones((rows, 2), np.float32)
    for _ in range(2):
        idx = (rng.zipf(1.3, size=n) % rows).astype(np.int64)
        g = rng.standard_normal((n, dim)).astype(np.float32)
        # WrappedLocalTensor keeps a POINTER, not the tensor: the operands must stay referenced until the call has been
        # issued (the first version passed temporaries; torch handed the freed index block to the gradient copy)
        idx_t, g_t = torch.from_numpy(idx).cuda(), torch.from_numpy(g).cuda()
        rwmb.EmbeddingGatherGradientApply(emb, wrap(idx_t), wrap(g_t), False, 0.02, env.get_env_fns(), stream)
        urows, ug = O.dedup_gradients(idx, g)
        O.optimizer_step("adam", w, urows, ug, 0.02, state=(m, v), b12=b12, beta1=0.85, weight_decay=0.01)
    torch.cuda.synchronize()
    got = local.cpu().numpy()
    if not np.allclose(got, w, rtol=1e-5, atol=1e-5):
        bad = np.argwhere(~np.isclose(got, w, rtol=1e-5, atol=1e-5))
        cnt = np.bincount(idx, minlength=rows)
        lines = ["row %d col %d: got %r oracle %r (dups of this row in the last step: %d)" % (r, c, got[r, c], w[r, c], cnt[r]) for r, c in bad[:8]]
        raise AssertionError("LazyAdam through the reference binding differs from the oracle: %d of %d elements in %d rows, max abs err %g\n%s"
                             % (bad.shape[0], got.size, np.unique(bad[:, 0]).size, np.abs(got - w).max(), "\n".join(lines)))
    # embedding gather through the reference binding
    gi = rng.integers(0, rows, size=777).astype(np.int64)
    out = torch.empty(777, dim, device="cuda")
    gi_t = torch.from_numpy(gi).cuda()
    rwmb.EmbeddingGatherForward(emb, wrap(gi_t), wrap(out), False, env.get_env_fns(), stream)
    torch.cuda.synchronize()
    assert torch.equal(out, local[torch.from_numpy(gi).cuda()])
    emb.destroy_embedding()
    opt.destroy_optimizer()

    # ---- CSR sampling + append_unique + self loops
    nodes = 1500
    deg = rng.integers(0, 70, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, nodes, size=int(row_ptr[-1])).astype(np.int64)
    rp = rwmb.create_wholememory_array(DT.DtInt64, nodes + 1, comm, rwmb.WholeMemoryMemoryType.MtChunked, rwmb.WholeMemoryMemoryLocation.MlDevice)
    cp = rwmb.create_wholememory_array(DT.DtInt64, col.size, comm, rwmb.WholeMemoryMemoryType.MtChunked, rwmb.WholeMemoryMemoryLocation.MlDevice)
    rp.get_local_tensor(from_dlpack, rwmb.WholeMemoryMemoryLocation.MlDevice, 0)[0].copy_(torch.from_numpy(row_ptr))
    cp.get_local_tensor(from_dlpack, rwmb.WholeMemoryMemoryLocation.MlDevice, 0)[0].copy_(torch.from_numpy(col))
    centers = rng.integers(0, nodes, size=600).astype(np.int64)
    offsets = torch.empty(601, dtype=torch.int32, device="cuda")
    c_dst, c_lid, c_gid = _Ctx(), _Ctx(), _Ctx()
    centers_t = torch.from_numpy(centers).cuda()  # referenced until the call returns (the env callbacks allocate from torch inside it)
    rwmb.csr_unweighted_sample_without_replacement(rp, cp, wrap(centers_t), 25, wrap(offsets), id(c_dst), id(c_lid),
                                                   id(c_gid), 4321, env.get_env_fns(), stream)
    torch.cuda.synchronize()
    eo, ed, el, eg = O.unweighted_sample(row_ptr, col, centers, 25, 4321)
    assert offsets.cpu().tolist() == eo.tolist() and c_dst.tensor.cpu().tolist() == ed.tolist()
    assert c_lid.tensor.cpu().tolist() == el.tolist() and c_gid.tensor.cpu().tolist() == eg.tolist()

    targets = torch.from_numpy(centers[:200].copy()).cuda().unique()
    c_unique = _Ctx()
    mapping = torch.empty(c_dst.tensor.shape[0], dtype=torch.int32, device="cuda")
    rwmb.append_unique(wrap(targets), wrap(c_dst.tensor), id(c_unique), wrap(mapping), env.get_env_fns(), stream)
    torch.cuda.synchronize()
    uniq = c_unique.tensor
    assert torch.equal(uniq[: targets.shape[0]], targets)
    assert torch.equal(uniq[mapping.long()], c_dst.tensor)
    assert uniq.unique().shape[0] == uniq.shape[0] == torch.cat([targets, c_dst.tensor]).unique().shape[0]

    r32 = offsets.clone()
    c32 = mapping.clone()
    orow = torch.empty_like(r32)
    ocol = torch.empty(c32.shape[0] + 600, dtype=torch.int32, device="cuda")
    rwmb.add_csr_self_loop(wrap(r32), wrap(c32), wrap(orow), wrap(ocol), stream)
    torch.cuda.synchronize()
    assert orow.cpu().tolist() == (r32.cpu() + torch.arange(601, dtype=torch.int32)).tolist()
    r_host, oc_host, c_host = orow.cpu().numpy(), ocol.cpu().numpy(), c32.cpu().numpy()
    old = r32.cpu().numpy()
    for i in (0, 1, 299, 599):  # row i: its own index first, then its old neighbours in order
        assert oc_host[r_host[i]] == i
        assert oc_host[r_host[i] + 1: r_host[i + 1]].tolist() == c_host[old[i]: old[i + 1]].tolist()

    rwmb.destroy_wholememory_tensor(rp)
    rwmb.destroy_wholememory_tensor(cp)
    rwmb.destroy_communicator(comm)
    rwmb.finalize()
    print("reference binding worker OK")


if __name__ == "__main__":
    main()
