"""The op wrappers of the reference's torch layer -- wholegraph_ops.py (both samplers and the two host random functions),
graph_ops.py, wholememory_ops.py, loaded unchanged through compat/ -- next to this repo's over one recording fake of the
binding.  The fake ops fill the caller's output contexts the way the library does (through the context object whose id they
were handed), so what comes back must be the same tuples in the same order for every combination of the optional outputs.
CPU only: "cuda" placements are redirected to the host."""
import ctypes
import os
import random
import types

import pytest
import torch

import wholegraph_b200.binding as wmb
import wholegraph_b200.torch.graph_ops as our_graph_ops
import wholegraph_b200.torch.wholegraph_ops as our_wg_ops
import wholegraph_b200.torch.wholememory_ops as our_wm_ops

REF_DIR = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not present")


class _HostTorch:
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def empty(*args, **kwargs):
        if str(kwargs.get("device", "")).startswith("cuda"):
            kwargs["device"] = "cpu"
        return torch.empty(*args, **kwargs)


def _ctx(address):
    return ctypes.cast(ctypes.c_void_p(address), ctypes.py_object).value if address else None


class _FakeWm:
    def __init__(self, n):
        self.shape, self.dtype = (n,), wmb.DtInt64

    def dim(self):
        return 1


@pytest.fixture()
def layers(monkeypatch):
    from compat_loader import load_reference_file
    ref = {n: load_reference_file(os.path.join(REF_DIR, n + ".py"), "_reference_" + n, package="pylibwholegraph.torch")
           for n in ("wholegraph_ops", "graph_ops", "wholememory_ops")}
    ours = {"wholegraph_ops": our_wg_ops, "graph_ops": our_graph_ops, "wholememory_ops": our_wm_ops}
    log = []

    def sample(kind):
        def op(*args):
            if kind == "weighted":
                rp, cp, wp, w_centers, k, w_offsets, dest, lid, gid, seed, env, stream = args
            else:
                rp, cp, w_centers, k, w_offsets, dest, lid, gid, seed, env, stream = args
            n = w_centers.t.shape[0]
            log.append((kind, tuple(w_centers.t.tolist()), k, seed, bool(lid), bool(gid), tuple(w_offsets.t.shape), w_offsets.t.dtype))
            w_offsets.t.copy_(torch.arange(n + 1, dtype=torch.int32) * 2)
            _ctx(dest).set_tensor(torch.arange(2 * n) + 100)
            if lid:
                _ctx(lid).set_tensor(torch.arange(2 * n, dtype=torch.int32) // 2)
            if gid:
                _ctx(gid).set_tensor(torch.arange(2 * n) + 1000)
        return op

    def append_unique(w_targets, w_neighbors, out_ctx, w_mapping, env, stream):
        log.append(("append_unique", tuple(w_targets.t.tolist()), tuple(w_neighbors.t.tolist()), w_mapping.t is not None))
        _ctx(out_ctx).set_tensor(torch.cat([w_targets.t, w_neighbors.t]).unique())
        if w_mapping.t is not None:
            w_mapping.t.copy_(torch.arange(w_neighbors.t.shape[0], dtype=torch.int32))

    def add_self_loop(w_row, w_col, w_out_row, w_out_col, stream):
        log.append(("add_csr_self_loop", tuple(w_out_row.t.shape), w_out_row.t.dtype, tuple(w_out_col.t.shape), w_out_col.t.dtype))
        w_out_row.t.zero_()
        w_out_col.t.zero_()

    def gather_op(t, w_idx, w_out, env, stream):
        log.append(("gather_op", tuple(w_idx.t.tolist()), tuple(w_out.t.shape), w_out.t.dtype, w_out.t.requires_grad))
        w_out.t.data.fill_(2.0)

    def scatter_op(w_in, w_idx, t, env, stream):
        log.append(("scatter_op", tuple(w_in.t.shape), tuple(w_idx.t.tolist())))

    monkeypatch.setattr(wmb, "csr_unweighted_sample_without_replacement", sample("unweighted"))
    monkeypatch.setattr(wmb, "csr_weighted_sample_without_replacement", sample("weighted"))
    monkeypatch.setattr(wmb, "append_unique", append_unique)
    monkeypatch.setattr(wmb, "add_csr_self_loop", add_self_loop)
    monkeypatch.setattr(wmb, "wholememory_gather_op", gather_op)
    monkeypatch.setattr(wmb, "wholememory_scatter_op", scatter_op)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    for group in (ours, ref):
        for mod in group.values():
            monkeypatch.setattr(mod, "torch", _HostTorch())
            monkeypatch.setattr(mod, "wrap_torch_tensor", lambda t: types.SimpleNamespace(t=t))
            monkeypatch.setattr(mod, "get_wholegraph_env_fns", lambda: 0)
            monkeypatch.setattr(mod, "get_stream", lambda: 0)
    return types.SimpleNamespace(ours=ours, ref=ref, log=log)


def _same(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dtype == b.dtype and torch.equal(a, b)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


def _both(layers, module, script):
    out = []
    for group in (layers.ours, layers.ref):
        del layers.log[:]
        random.seed(1234)  # an unseeded sampler call draws its seed from `random`
        try:
            result = ("ok", script(group[module]))
        except Exception as e:
            result = ("raises", type(e).__name__)
        out.append((result, list(layers.log)))
    assert _same(out[0], out[1]), out
    return out[0]


def test_sampler_wrappers(layers):
    rp, cp, wp = _FakeWm(11), _FakeWm(50), _FakeWm(50)
    centers = torch.tensor([3, 1, 4, 1, 5])
    for need_lid in (False, True):
        for need_gid in (False, True):
            for seed in (None, 77, (1 << 64) - 1):
                result, log = _both(layers, "wholegraph_ops", lambda m: m.unweighted_sample_without_replacement(
                    rp, cp, centers, 25, seed, need_lid, need_gid))
                assert result[0] == "ok" and len(result[1]) == 2 + need_lid + need_gid and log[0][3] is not None
                _both(layers, "wholegraph_ops", lambda m: m.weighted_sample_without_replacement(
                    rp, cp, wp, centers, 10, random_seed=seed, need_center_local_output=need_lid, need_edge_output=need_gid))
    _both(layers, "wholegraph_ops", lambda m: m.unweighted_sample_without_replacement(rp, cp, centers.reshape(5, 1), 25))   # 2-D centers
    _both(layers, "wholegraph_ops", lambda m: m.weighted_sample_without_replacement(rp, cp, _FakeWm(49), centers, 25))     # weights != edges


def test_host_random_wrappers_call_the_library(layers, monkeypatch):
    """No fake here: both layers call this library's host functions."""
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    for group in (layers.ours, layers.ref):
        monkeypatch.setattr(group["wholegraph_ops"], "wrap_torch_tensor", wrap_torch_tensor)
    for fn, args in (("generate_random_positive_int_cpu", (42, 3, 9)), ("generate_exponential_distribution_negative_float_cpu", (42, 3, 9))):
        a = getattr(layers.ours["wholegraph_ops"], fn)(*args)
        b = getattr(layers.ref["wholegraph_ops"], fn)(*args)
        assert a.dtype == b.dtype and torch.equal(a, b) and a.shape[0] == 9


def test_graph_and_gather_scatter_wrappers(layers):
    t, n = torch.tensor([3, 11, 2, 10], dtype=torch.int32), torch.tensor([4, 5, 2, 11, 6, 9, 10, 5], dtype=torch.int32)
    for need in (False, True):
        _both(layers, "graph_ops", lambda m: m.append_unique(t, n, need_neighbor_raw_to_unique=need))
    _both(layers, "graph_ops", lambda m: m.append_unique(t.reshape(2, 2), n))
    row, col = torch.tensor([0, 2, 3], dtype=torch.int32), torch.tensor([1, 0, 1], dtype=torch.int32)
    _both(layers, "graph_ops", lambda m: m.add_csr_self_loop(row, col))
    table = types.SimpleNamespace(shape=(40, 6), dtype=wmb.DtFloat)
    idx = torch.tensor([1, 2, 3])
    _both(layers, "wholememory_ops", lambda m: m.wholememory_gather_forward_functor(table, idx))
    _both(layers, "wholememory_ops", lambda m: m.wholememory_gather_forward_functor(table, idx, True, torch.float16))
    _both(layers, "wholememory_ops", lambda m: m.wholememory_gather_forward_functor(table, idx.to(torch.int16)))
    _both(layers, "wholememory_ops", lambda m: m.wholememory_scatter_functor(torch.ones(3, 6), idx, table))
    _both(layers, "wholememory_ops", lambda m: m.wholememory_scatter_functor(torch.ones(3, 6), idx.reshape(3, 1), table))
