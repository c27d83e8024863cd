"""The factory side of pylibwholegraph/torch/embedding.py (loaded unchanged through compat/) next to this repo's, over one
recording fake of the binding: create_embedding, create_embedding_from_filelist, create_wholememory_cache_policy,
create_builtin_cache_policy, create_wholememory_optimizer, the destroy functions and save / load must make the same
binding calls with the same arguments (valid and invalid input).  HIERARCHY / NVSHMEM corners are outside this build's scope
and not exercised.  CPU only."""
import itertools
import os
import types

import pytest
import torch

import wholegraph_b200.binding as wmb
from wholegraph_b200.torch import embedding as our_mod
from wholegraph_b200.torch import tensor as tensor_mod

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/embedding.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


class _Comm:
    distributed_backend = "nccl"

    def __init__(self, log, name):
        self.log, self.wmb_comm, self.name = log, "wmb:" + name, name

    def barrier(self):
        self.log.append(("barrier", self.name))


class _FakeTensor:
    def __init__(self, log, name, shape):
        self.log, self.name, self.shape, self.dtype = log, name, shape, wmb.DtFloat

    def dim(self):
        return 2

    def get_wholememory_handle(self):
        return types.SimpleNamespace(get_communicator=lambda: types.SimpleNamespace(get_rank=lambda: 1, get_size=lambda: 4))

    def get_local_tensor(self, *args):
        self.log.append((self.name, "get_local_tensor"))
        return torch.zeros(5, 4), 10

    def from_filelist(self, filelist, round_robin_size=0):
        self.log.append((self.name, "from_filelist", tuple(filelist), round_robin_size))

    def to_file(self, filename):
        self.log.append((self.name, "to_file", filename))


class _FakeEmbedding:
    def __init__(self, log, shape):
        self.log, self.shape = log, shape

    def get_embedding_tensor(self):
        self.log.append(("embedding", "get_embedding_tensor"))
        return _FakeTensor(self.log, "table", self.shape)

    def get_optimizer_state_names(self):
        return ["m", "v", "beta12t"]

    def get_optimizer_state(self, name):
        self.log.append(("embedding", "get_optimizer_state", name))
        return _FakeTensor(self.log, "state:" + name, self.shape)

    def destroy_embedding(self):
        self.log.append(("embedding", "destroy"))


@pytest.fixture()
def layers(monkeypatch):
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF, "_reference_embedding_factories", package="pylibwholegraph.torch")
    log = []

    class FakePolicy:
        def create_policy(self, comm, memory_type, location, access, ratio):
            log.append(("policy.create", comm, int(memory_type), int(location), int(access), float(ratio)))

        def destroy_policy(self):
            log.append(("policy.destroy",))

    class FakeOptimizer:
        def create_optimizer(self, optimizer_type, params):
            log.append(("optimizer.create", int(optimizer_type), dict(params)))

        def add_embedding(self, handle):
            log.append(("optimizer.add_embedding", type(handle).__name__))

        def destroy_optimizer(self):
            log.append(("optimizer.destroy",))

    def create_embedding(desc, comm, memory_type, location, policy, embedding_entry_partition=None, user_defined_sms=-1, round_robin_size=0):
        log.append(("create_embedding", tuple(desc.shape), tuple(desc.stride()), int(desc.dtype), comm, int(memory_type), int(location),
                    type(policy).__name__, None if embedding_entry_partition is None else tuple(embedding_entry_partition),
                    user_defined_sms, round_robin_size))
        return _FakeEmbedding(log, tuple(desc.shape))

    monkeypatch.setattr(wmb, "create_embedding", create_embedding)
    monkeypatch.setattr(wmb, "create_non_cache_policy", lambda: "no-cache")
    monkeypatch.setattr(wmb, "WholeMemoryCachePolicy", FakePolicy)
    monkeypatch.setattr(wmb, "WholeMemoryOptimizer", FakeOptimizer)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.nn.init, "xavier_uniform_", lambda t: log.append(("xavier", tuple(t.shape))))
    # the communicator getters of both layers are this repo's comm module (aliased): give it cached fakes
    from wholegraph_b200.torch import comm as comm_mod
    comms = {n: _Comm(log, n) for n in ("global", "node", "device")}
    monkeypatch.setattr(comm_mod, "get_global_communicator", lambda *a: comms["global"])
    for mod in (our_mod, ref_mod):
        monkeypatch.setattr(mod, "get_global_communicator", lambda *a: comms["global"])
        monkeypatch.setattr(mod, "get_local_node_communicator", lambda: comms["node"])
        monkeypatch.setattr(mod, "get_local_device_communicator", lambda: comms["device"])
    monkeypatch.setattr(our_mod, "_BUILTIN_CACHE_COMM", {"all_devices": lambda: comms["global"], "local_node": lambda: comms["node"],
                                                         "local_device": lambda: comms["device"]})  # ours looks the getter up in a table
    monkeypatch.setattr(tensor_mod, "wholememory_dtype_to_torch_dtype", lambda d: torch.float32)
    return types.SimpleNamespace(ours=our_mod, ref=ref_mod, log=log, comm=_Comm(log, "user"))


def _outcome(fn):
    try:
        r = fn()
        return ("ok", type(r).__name__)
    except Exception as e:
        return ("raises", type(e).__name__)


def _both(layers, script):
    out = []
    for mod in (layers.ours, layers.ref):
        del layers.log[:]
        out.append((_outcome(lambda: script(mod)), list(layers.log)))
    assert out[0] == out[1], out
    return out[0]


def test_create_embedding_variants(layers, tmp_path, capsys):
    policy_holder = types.SimpleNamespace(wmb_cache_policy=types.SimpleNamespace())
    for (mt, loc), dtype, partition, cache, rr, sms, init in itertools.product(
            [("chunked", "cuda"), ("distributed", "cuda"), ("continuous", "cpu"), ("nonsense", "cuda")], [torch.float32, torch.float16],
            [None, [3, 7]], [None, policy_holder], [0, 64], [-1, 32], [False, True]):
        _both(layers, lambda m: m.create_embedding(layers.comm, mt, loc, dtype, [10, 4], cache_policy=cache, embedding_entry_partition=partition,
                                                   random_init=init, gather_sms=sms, round_robin_size=rr))
    _both(layers, lambda m: m.create_embedding(layers.comm, "chunked", "cuda", torch.float32, [10]))          # not 2-D
    a, b, odd = tmp_path / "a.bin", tmp_path / "b.bin", tmp_path / "odd.bin"
    a.write_bytes(b"\0" * (4 * 4 * 10))
    b.write_bytes(b"\0" * (4 * 4 * 2))
    odd.write_bytes(b"\0" * 30)
    for files, last, partition, rr in ((str(a), 4, None, 0), ([str(a), str(b)], 4, [5, 7], 16), ([str(a), str(odd)], 4, None, 0),
                                       ([str(a)], 0, None, 0), ([str(tmp_path / "none")], 4, None, 0)):
        _both(layers, lambda m: m.create_embedding_from_filelist(layers.comm, "chunked", "cuda", files, torch.float32, last,
                                                                 embedding_entry_partition=partition, round_robin_size=rr))
    capsys.readouterr()


def test_cache_policy_factories(layers, capsys):
    for kw in ({}, {"memory_type": "continuous", "memory_location": "cpu", "access_type": "readwrite", "ratio": 0.1}, {"memory_type": "x"},
               {"access_type": "sometimes"}, {"memory_location": "moon"}):
        _both(layers, lambda m: m.create_wholememory_cache_policy(layers.comm, **kw))
    for builtin, emb_type, emb_loc, access, ratio, cache_type, cache_loc in itertools.product(
            ["none", "all_devices", "local_node", "local_device", "other"], ["continuous", "chunked", "distributed", "bad"], ["cpu", "cuda", "moon"],
            ["readonly", "readwrite"], [0.25], ["", "chunked"], ["", "cpu", "cuda", "moon"]):
        _both(layers, lambda m: m.create_builtin_cache_policy(builtin, emb_type, emb_loc, access, ratio, cache_memory_type=cache_type,
                                                              cache_memory_location=cache_loc))
    _both(layers, lambda m: m.destroy_wholememory_cache_policy(m.create_wholememory_cache_policy(layers.comm)))
    capsys.readouterr()


def test_optimizer_factory_save_load_destroy(layers):
    def make(m):
        return m.create_embedding(layers.comm, "chunked", "cuda", torch.float32, [10, 4])

    for kind, params in (("sgd", {}), ("adam", {"beta1": 0.8, "adam_w": 1.0}), ("adagrad", {"epsilon": 1e-6}), ("rmsprop", {}), ("lion", {})):
        _both(layers, lambda m: m.create_wholememory_optimizer(make(m), kind, params))
        _both(layers, lambda m: m.create_wholememory_optimizer([make(m), make(m)], kind, params))
    _both(layers, lambda m: m.destroy_wholememory_optimizer(m.create_wholememory_optimizer(make(m), "sgd", {})))
    _both(layers, lambda m: m.destroy_embedding(make(m)))
    _both(layers, lambda m: make(m).save("ckpt"))
    _both(layers, lambda m: make(m).load("ckpt"))
    _both(layers, lambda m: make(m).load("ckpt", ignore_embedding=True, part_count=3))

    def states_twice(m):
        e = make(m)
        assert e.get_optimizer_state("m") is e.get_optimizer_state("m") and e.get_embedding_tensor() is e.get_embedding_tensor()
        return e.get_optimizer_state_names()
    _both(layers, states_twice)
