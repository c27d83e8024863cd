"""The user-facing training path on the GPU: WholeMemoryEmbeddingModule + torch autograd + WholeMemoryOptimizer.step
(reference pylibwholegraph/torch/embedding.py:33-70, :213-243, :537-587), three steps of a toy regression with duplicate
indices, for each optimizer; weights compared with the oracle's dedup + optimizer restatement at the reference's 1e-5.

(File name sorts last on purpose: added without a GPU at hand; tests/test_torch_training_flow.py covers the control flow on CPU.)"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("kind,params", [("sgd", {}), ("adam", {}), ("adagrad", {"epsilon": 1e-6}), ("rmsprop", {"alpha": 0.9})])
def test_module_autograd_optimizer_step(kind, params):
    import gpu_utils as G
    import wholegraph_b200.torch as wgth
    comm = wgth.WholeMemoryCommunicator(G.single_comm())
    rows, dim, n, lr = 500, 48, 300, 0.05
    rng = np.random.default_rng(17)
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, dim])
    opt = wgth.create_wholememory_optimizer(emb, kind, params, global_comm=comm)
    local, first = emb.get_embedding_tensor().get_local_tensor()
    w = rng.standard_normal((rows, dim)).astype(np.float32)
    local.copy_(torch.from_numpy(w))
    module = wgth.WholeMemoryEmbeddingModule(emb)
    module.train()
    m, v, b12 = np.zeros_like(w), np.zeros_like(w), np.ones((rows, 2), np.float32)
    try:
        for step in range(3):
            idx = (rng.zipf(1.3, size=n) % rows).astype(np.int64)
            coef = rng.standard_normal((n, dim)).astype(np.float32)
            out = module(torch.from_numpy(idx).cuda())
            assert out.requires_grad and tuple(out.shape) == (n, dim)
            assert torch.equal(out.detach().cpu(), torch.from_numpy(w[idx])) or np.allclose(out.detach().cpu().numpy(), w[idx], rtol=1e-5, atol=1e-5)
            loss = (out * torch.from_numpy(coef).cuda()).sum()  # d loss / d out = coef
            loss.backward()
            opt.step(lr)
            torch.cuda.synchronize()
            urows, ug = O.dedup_gradients(idx, coef)
            kw = dict(weight_decay=params.get("weight_decay", 0.0), epsilon=params.get("epsilon", 1e-8))
            if kind == "adam":
                O.optimizer_step("adam", w, urows, ug, lr, state=(m, v), b12=b12, **kw)
            elif kind == "sgd":
                O.optimizer_step("sgd", w, urows, ug, lr, weight_decay=kw["weight_decay"])
            elif kind == "adagrad":
                O.optimizer_step("adagrad", w, urows, ug, lr, state=m, **kw)
            else:
                O.optimizer_step("rmsprop", w, urows, ug, lr, state=m, alpha=params.get("alpha", 0.99), **kw)
            assert np.allclose(local.cpu().numpy(), w, rtol=1e-5, atol=1e-5), "%s step %d" % (kind, step)
            assert not emb.need_apply and emb.sparse_indices == []
        module.eval()
        out = module(torch.arange(10, device="cuda"))
        assert np.allclose(out.detach().cpu().numpy(), w[:10], rtol=1e-5, atol=1e-5)
    finally:
        wgth.destroy_wholememory_optimizer(opt)
        wgth.destroy_embedding(emb)
