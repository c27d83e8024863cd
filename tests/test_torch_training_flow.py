"""Control flow of the torch-level training path -- WholeMemoryEmbeddingModule -> EmbeddingLookupFn (autograd) ->
WholeMemoryEmbedding.add_gradients -> WholeMemoryOptimizer.step -> one gradient-apply call per embedding -- with the C
library replaced by an in-memory fake, so it runs without a GPU.  What is checked is the Python layer's contract with the
binding (reference pylibwholegraph/torch/embedding.py:33-70, :213-243, :280-335, :537-555): which binding calls are made,
with which tensors, in which order; that gradients reach the embedding only in training mode and only when an optimizer is
attached; that need_apply / the pending lists are cleared after a step; that the barrier follows the applies.
The arithmetic itself belongs to the GPU tests (tests/test_zz_training_autograd_gpu.py)."""
import types

import pytest
import torch

import wholegraph_b200.torch as wgth
from wholegraph_b200.torch import embedding as emb_mod
from wholegraph_b200.torch import tensor as tensor_mod


class _FakeTensorHandle:
    """Stands in for wmb.PyWholeMemoryTensor over a host torch tensor."""

    def __init__(self, data):
        self.data = data
        self.dtype = "fake-fp32"
        self.shape = tuple(data.shape)

    def dim(self):
        return self.data.dim()


class _FakeEmbeddingHandle:
    def __init__(self, rows, cols):
        self.table = _FakeTensorHandle(torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols) / 10.0)

    def get_embedding_tensor(self):
        return self.table

    def get_optimizer_state_names(self):
        return []


class _FakeOptimizerHandle:
    def __init__(self):
        self.added = []

    def add_embedding(self, handle):
        self.added.append(handle)


class _FakeComm:
    def __init__(self, log):
        self.log = log

    def barrier(self):
        self.log.append(("barrier",))


@pytest.fixture()
def fake(monkeypatch):
    log = []

    def gather_forward(handle, w_idx, w_out, adjust_cache, env, stream):
        idx, out = w_idx._keepalive, w_out._keepalive
        out.data.copy_(handle.table.data[idx])
        log.append(("gather", tuple(idx.tolist()), adjust_cache))

    def gradient_apply(handle, w_idx, w_grads, adjust_cache, lr, env, stream):
        idx, g = w_idx._keepalive, w_grads._keepalive
        handle.table.data.index_add_(0, idx, -lr * g)  # plain SGD with duplicate accumulation: enough to see the data flow
        log.append(("apply", tuple(idx.tolist()), float(lr)))

    def wrap(t):
        return types.SimpleNamespace(_keepalive=t)

    for mod in (emb_mod, tensor_mod):
        monkeypatch.setattr(mod, "wrap_torch_tensor", wrap)
        monkeypatch.setattr(mod, "get_wholegraph_env_fns", lambda: 0)
        monkeypatch.setattr(mod, "get_stream", lambda: 0)
        monkeypatch.setattr(mod, "current_output_device", lambda: "cpu")
        monkeypatch.setattr(mod, "wholememory_dtype_to_torch_dtype", lambda d: torch.float32, raising=False)
    monkeypatch.setattr(emb_mod.wmb, "EmbeddingGatherForward", gather_forward)
    monkeypatch.setattr(emb_mod.wmb, "EmbeddingGatherGradientApply", gradient_apply)
    monkeypatch.setattr(emb_mod.wmb, "WholeMemoryOptimizer", _FakeOptimizerHandle)
    return log


def _make(fake_log, rows=8, cols=4, with_optimizer=True):
    emb = wgth.WholeMemoryEmbedding(_FakeEmbeddingHandle(rows, cols), None)
    opt = None
    if with_optimizer:
        opt = wgth.WholeMemoryOptimizer(_FakeComm(fake_log))
        opt.add_embedding(emb)
    return emb, opt


def test_training_step_through_autograd(fake):
    emb, opt = _make(fake)
    assert emb.need_grad() and emb.dummy_input.requires_grad and opt.wmb_opt.added == [emb.wmb_embedding]
    with pytest.raises(ValueError):
        opt.add_embedding(emb)  # an embedding takes one optimizer, once
    module = wgth.WholeMemoryEmbeddingModule(emb)
    module.train()
    before = emb.wmb_embedding.table.data.clone()
    idx = torch.tensor([3, 5, 3])
    rows = module(idx)
    assert rows.requires_grad and emb.need_apply and torch.equal(rows.detach(), before[idx])
    weights = torch.tensor([[1.0], [2.0], [4.0]])
    (rows * weights).sum().backward()
    assert len(emb.sparse_indices) == 1 and torch.equal(emb.sparse_indices[0], idx)
    assert torch.equal(emb.sparse_grads[0], weights.expand(3, 4))
    # a second micro-batch before the step: both are applied in ONE call, in order
    rows2 = module(torch.tensor([0]))
    rows2.sum().backward()
    opt.step(0.5)
    assert [e[0] for e in fake] == ["gather", "gather", "apply", "barrier"]
    assert fake[2] == ("apply", (3, 5, 3, 0), 0.5)
    assert emb.sparse_indices == [] and emb.sparse_grads == [] and not emb.need_apply
    after = emb.wmb_embedding.table.data
    assert torch.equal(after[3], before[3] - 0.5 * (1.0 + 4.0)) and torch.equal(after[5], before[5] - 0.5 * 2.0)
    assert torch.equal(after[0], before[0] - 0.5) and torch.equal(after[1], before[1])
    # a step with nothing pending only barriers
    opt.step(0.5)
    assert [e[0] for e in fake][-1] == "barrier" and [e[0] for e in fake].count("apply") == 1


def test_no_gradients_in_eval_mode_or_without_optimizer(fake):
    emb, opt = _make(fake)
    module = wgth.WholeMemoryEmbeddingModule(emb)
    module.eval()
    rows = module(torch.tensor([1, 2]))
    # (the autograd node still exists because dummy_input requires grad -- as in the reference -- but nothing is recorded)
    assert not emb.need_apply and emb.sparse_indices == [] and torch.equal(rows.detach(), emb.wmb_embedding.table.data[1:3])
    emb2, _ = _make(fake, with_optimizer=False)
    module2 = wgth.WholeMemoryEmbeddingModule(emb2)
    module2.train()
    rows = module2(torch.tensor([1, 2]))
    # without an optimizer nothing can reach backward: dummy_input does not require grad, so autograd drops the node and the
    # rows come back not requiring grad.  (need_grad() is True for any live embedding and need_apply gets set, exactly as in
    # the reference -- embedding.py:276-277, :300-301 -- but no optimizer ever looks at this embedding.)
    assert emb2.need_grad() and not rows.requires_grad and emb2.sparse_indices == []


def test_force_dtype_and_direct_gather(fake):
    emb, _ = _make(fake)
    rows = emb.gather(torch.tensor([7]), force_dtype=torch.float64)
    assert rows.dtype == torch.float64 and not rows.requires_grad
    t = emb.get_embedding_tensor()
    assert t is emb.get_embedding_tensor() and t.shape == (8, 4)
