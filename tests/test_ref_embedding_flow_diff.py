"""The reference's embedding.py classes (WholeMemoryEmbedding, EmbeddingLookupFn, WholeMemoryEmbeddingModule,
WholeMemoryOptimizer -- loaded unchanged through compat/) next to this repo's, over the same in-memory fake of the binding:
the same user script must produce the same sequence of binding calls, the same tensors and the same table.
CPU only (the reference allocates its outputs on "cuda:<n>"; the module's `torch` global is proxied so that those land on the
host).  Deliberate differences are listed in DIFFERENCES below and in DESIGN.md section 7."""
import os
import types

import pytest
import torch

import wholegraph_b200.torch as wgth
from wholegraph_b200.torch import embedding as our_mod

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/embedding.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


class _HostTorch:
    """`torch` as the reference module sees it, with cuda placement redirected to the host."""

    class _Cuda:
        @staticmethod
        def current_device():
            return 0

    cuda = _Cuda()

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def empty(*args, **kwargs):
        if str(kwargs.get("device", "")).startswith("cuda"):
            kwargs["device"] = "cpu"
        return torch.empty(*args, **kwargs)


class _Table:
    def __init__(self, rows, cols):
        self.data = torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols) / 10.0
        self.dtype, self.shape = "fake-fp32", (rows, cols)

    def dim(self):
        return 2


class _EmbHandle:
    def __init__(self, rows, cols):
        self.table = _Table(rows, cols)

    def get_embedding_tensor(self):
        return self.table

    def get_optimizer_state_names(self):
        return []


class _OptHandle:
    def __init__(self):
        self.added = []

    def add_embedding(self, handle):
        self.added.append(handle)


class _Comm:
    def __init__(self, log):
        self.log = log

    def barrier(self):
        self.log.append(("barrier",))


def _install_fake(monkeypatch, mod, log):
    def gather_forward(handle, w_idx, w_out, adjust_cache, env, stream):
        w_out.t.data.copy_(handle.table.data[w_idx.t])
        log.append(("gather", tuple(w_idx.t.tolist()), bool(adjust_cache), w_out.t.dtype, w_out.t.requires_grad))

    def gradient_apply(handle, w_idx, w_grads, adjust_cache, lr, env, stream):
        handle.table.data.index_add_(0, w_idx.t, -lr * w_grads.t)
        log.append(("apply", tuple(w_idx.t.tolist()), tuple(w_grads.t.flatten().tolist()), float(lr)))

    monkeypatch.setattr(mod, "wrap_torch_tensor", lambda t: types.SimpleNamespace(t=t))
    monkeypatch.setattr(mod, "get_wholegraph_env_fns", lambda: 0)
    monkeypatch.setattr(mod, "get_stream", lambda *a: 0)
    monkeypatch.setattr(mod.wmb, "EmbeddingGatherForward", gather_forward)
    monkeypatch.setattr(mod.wmb, "EmbeddingGatherGradientApply", gradient_apply)
    monkeypatch.setattr(mod.wmb, "WholeMemoryOptimizer", _OptHandle)


def _script(mod, log, with_optimizer):
    """A user's training loop: two micro-batches, a step, an eval lookup, a forced dtype, an empty step."""
    emb = mod.WholeMemoryEmbedding(_EmbHandle(8, 4), None)
    observed = []
    if with_optimizer:
        opt = mod.WholeMemoryOptimizer(_Comm(log))
        opt.add_embedding(emb)
    module = mod.WholeMemoryEmbeddingModule(emb)
    module.train()
    rows = module(torch.tensor([3, 5, 3]))
    observed.append(("train rows", rows.detach().clone(), rows.requires_grad, bool(emb.need_apply)))
    if rows.requires_grad:
        (rows * torch.tensor([[1.0], [2.0], [4.0]])).sum().backward()
    rows2 = module(torch.tensor([0]))
    if rows2.requires_grad:
        rows2.sum().backward()
    observed.append(("pending", [i.tolist() for i in emb.sparse_indices], [g.tolist() for g in emb.sparse_grads]))
    if with_optimizer:
        opt.step(0.5)
        observed.append(("after step", list(emb.sparse_indices), list(emb.sparse_grads), bool(emb.need_apply)))
        opt.step(0.25)  # nothing pending
    module.eval()
    rows3 = module(torch.tensor([1, 2]))
    observed.append(("eval rows", rows3.detach().clone(), [i.tolist() for i in emb.sparse_indices]))
    half = module(torch.tensor([6]), force_dtype=torch.float16)
    observed.append(("forced dtype", half.dtype, half.detach().clone()))
    observed.append(("table", emb.wmb_embedding.table.data.clone()))
    return observed


def _same(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dtype == b.dtype and torch.equal(a, b)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


@pytest.mark.parametrize("with_optimizer", [True, False])
def test_same_script_same_binding_calls(monkeypatch, with_optimizer):
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF, "_reference_embedding", package="pylibwholegraph.torch")
    monkeypatch.setattr(ref_mod, "torch", _HostTorch())
    monkeypatch.setattr(our_mod, "current_output_device", lambda: "cpu")
    for mod in (our_mod, ref_mod):
        assert mod.wmb is our_mod.wmb  # one binding module, patched once more below (idempotent)
    # the lazily created WholeMemoryTensor of both layers reads dtype / shape from the fake table
    from wholegraph_b200.torch import tensor as tensor_mod
    monkeypatch.setattr(tensor_mod, "wholememory_dtype_to_torch_dtype", lambda d: torch.float32)
    logs = {}
    results = {}
    for name, mod in (("ours", our_mod), ("reference", ref_mod)):
        logs[name] = []
        _install_fake(monkeypatch, mod, logs[name])
        results[name] = _script(mod, logs[name], with_optimizer)
    assert len(logs["ours"]) == len(logs["reference"]), (logs["ours"], logs["reference"])
    for a, b in zip(logs["ours"], logs["reference"]):
        assert _same(a, b), (a, b)
    for a, b in zip(results["ours"], results["reference"]):
        assert _same(a, b), (a, b)
