"""wholememory_embedding_gather (csrc/embedding.cpp; reference noncached_embedding::gather, cpp/src/wholememory/embedding.cpp:553-562)
on embeddings whose row stride is PADDED past the embedding dim (the reference pads rows to 16 bytes, embedding.cpp:43-50):
D = 127 fp32 -> stride 128, D = 100 fp16 -> stride 104.  The gathered rows must be the first D columns of the padded rows,
bit for bit, for every memory type, fp32 and fp16 outputs, int32 and int64 ids, including negative ids."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mem_type", ["continuous", "chunked", "distributed"])
@pytest.mark.parametrize("emb_dt,dim,padded", [(O.DT_FLOAT, 127, 128), (O.DT_HALF, 100, 104), (O.DT_FLOAT, 1, 4), (O.DT_FLOAT, 256, 256)])
@pytest.mark.parametrize("idx_dtype", [np.int32, np.int64])
def test_embedding_gather_on_padded_rows(mem_type, emb_dt, dim, padded, idx_dtype):
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    comm = G.single_comm()
    rows, n = 4099, 3000
    rng = np.random.default_rng(dim + len(mem_type))
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(G.WM_OF[emb_dt])
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, G.MT[mem_type], wmb.MlDevice, wmb.create_non_cache_policy())
    try:
        wt = emb.get_embedding_tensor()
        assert tuple(wt.shape) == (rows, dim) and wt.stride()[0] == padded, (wt.shape, wt.stride())
        local, off = wt.get_local_tensor(wmb.MlDevice, 0)  # [rows, dim] view with row stride `padded`
        assert off == 0 and tuple(local.shape) == (rows, dim) and local.stride(0) == padded
        host = G.random_table(rng, emb_dt, rows, dim)
        local.copy_(G.np_to_torch(host, emb_dt).cuda())
        idx = rng.integers(0, rows, size=n).astype(idx_dtype)
        idx[::9] = -1
        for out_dt in (O.DT_FLOAT, O.DT_HALF):
            sentinel = G.random_table(rng, out_dt, n, dim)
            out_t = G.np_to_torch(sentinel.copy(), out_dt).cuda()
            wmb.EmbeddingGatherForward(emb, wrap_torch_tensor(G.idx_to_cuda(idx)), wrap_torch_tensor(out_t), False, get_wholegraph_env_fns(),
                                       get_stream())
            torch.cuda.synchronize()
            exp = sentinel.copy()
            O.gather(host, emb_dt, idx, out_dt, out=exp)
            assert G.torch_to_np(out_t, out_dt).tobytes() == exp.tobytes(), "embedding gather %s -> %s" % (emb_dt, out_dt)
    finally:
        emb.destroy_embedding()
