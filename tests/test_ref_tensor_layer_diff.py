"""pylibwholegraph/torch/tensor.py of the reference (loaded unchanged through compat/) next to wholegraph_b200/torch/tensor.py
over one recording fake of the binding: every method of WholeMemoryTensor and the two factory functions must make the same
binding calls with the same arguments and hand back the same results.  CPU only.

Known difference, on the reference's side: its get_all_chunked_tensor names a method that does not exist
(`get_global_tensorget_all_chunked_tensor`, tensor.py:154-165) and raises AttributeError; this repo's works."""
import os
import types

import pytest
import torch

import wholegraph_b200.binding as wmb
from wholegraph_b200.torch import tensor as our_mod

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/tensor.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


class _Handle:
    def get_communicator(self):
        return types.SimpleNamespace(get_rank=lambda: 2, get_size=lambda: 5)


class _FakeWmbTensor:
    """Records what the torch layer asks of a PyWholeMemoryTensor; DLPack importers in first position are dropped."""

    def __init__(self, log, shape=(40, 6), name="root"):
        self.log, self.shape, self.name, self.dtype = log, shape, name, wmb.DtFloat

    def _rec(self, what, *args):
        args = args[1:] if args and callable(args[0]) else args
        self.log.append((self.name, what) + tuple(int(a) if isinstance(a, (int, wmb.WholeMemoryMemoryLocation)) else a for a in args))

    def dim(self):
        return len(self.shape)

    def stride(self):
        return (self.shape[1], 1) if len(self.shape) == 2 else (1,)

    def storage_offset(self):
        return 0

    def get_wholememory_handle(self):
        return _Handle()

    def get_sub_tensor(self, starts, ends):
        self._rec("get_sub_tensor", tuple(starts), tuple(ends))
        return _FakeWmbTensor(self.log, self.shape, self.name + ".sub")

    def get_local_tensor(self, *args):
        self._rec("get_local_tensor", *args)
        return torch.zeros(3), 7

    def get_global_tensor(self, *args):
        self._rec("get_global_tensor", *args)
        return torch.zeros(4)

    def get_all_chunked_tensor(self, *args):
        self._rec("get_all_chunked_tensor", *args)
        return [torch.zeros(1)], [0]

    def from_filelist(self, filelist, round_robin_size=0):
        self._rec("from_filelist", tuple(filelist), round_robin_size)

    def to_file(self, filename):
        self._rec("to_file", filename)


class _HostTorch:
    cuda = types.SimpleNamespace(current_device=lambda: 3)

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def empty(*args, **kwargs):
        if str(kwargs.get("device", "")).startswith("cuda"):
            kwargs["device"] = "cpu"
        return torch.empty(*args, **kwargs)


@pytest.fixture()
def layers(monkeypatch):
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF, "_reference_tensor", package="pylibwholegraph.torch")
    log = []

    def gather_op(t, w_idx, w_out, env, stream):
        log.append((t.name, "gather_op", tuple(w_idx.t.tolist()), tuple(w_out.t.shape), w_out.t.dtype, w_out.t.requires_grad))
        w_out.t.fill_(1.5)

    def scatter_op(w_in, w_idx, t, env, stream):
        log.append((t.name, "scatter_op", tuple(w_in.t.shape), tuple(w_idx.t.tolist())))

    def create_tensor(td, comm, memory_type, location, partition):
        log.append(("create", tuple(td.shape), tuple(td.stride()), int(td.dtype), comm, int(memory_type), int(location),
                    None if partition is None else tuple(partition)))
        return _FakeWmbTensor(log, tuple(td.shape), "created")

    def destroy_tensor(t):
        log.append((t.name, "destroy"))

    monkeypatch.setattr(wmb, "wholememory_gather_op", gather_op)
    monkeypatch.setattr(wmb, "wholememory_scatter_op", scatter_op)
    monkeypatch.setattr(wmb, "create_wholememory_tensor", create_tensor)
    monkeypatch.setattr(wmb, "destroy_wholememory_tensor", destroy_tensor)
    monkeypatch.setattr(ref_mod, "torch", _HostTorch())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 3)
    monkeypatch.setattr(our_mod, "current_output_device", lambda: "cpu")
    for mod in (our_mod, ref_mod):
        monkeypatch.setattr(mod, "wrap_torch_tensor", lambda t: types.SimpleNamespace(t=t))
        monkeypatch.setattr(mod, "get_wholegraph_env_fns", lambda: 0)
        monkeypatch.setattr(mod, "get_stream", lambda: 0)
    return types.SimpleNamespace(ours=our_mod, ref=ref_mod, log=log)


def _same(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dtype == b.dtype and torch.equal(a, b)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


def _outcome(fn):
    try:
        return ("ok", fn())
    except Exception as e:
        return ("raises", type(e).__name__)


def _both(layers, script):
    out = []
    for mod in (layers.ours, layers.ref):
        del layers.log[:]
        result = _outcome(lambda: script(mod))
        out.append((result, list(layers.log)))
    assert _same(out[0], out[1]), out
    return out[0]


def test_tensor_methods_make_the_same_binding_calls(layers, tmp_path):
    idx = torch.tensor([4, 0, 39])
    scripts = {
        "describe": lambda m: (m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).dtype, m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).shape,
                               m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).dim(), m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).stride(),
                               m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).storage_offset()),
        "gather": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).gather(idx),
        "gather forced dtype": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).gather(idx, force_dtype=torch.float16),
        "gather 2-D indices": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).gather(idx.reshape(3, 1)),
        "scatter": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).scatter(torch.ones(3, 6), idx),
        "scatter wrong width": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).scatter(torch.ones(3, 5), idx),
        "scatter wrong count": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).scatter(torch.ones(2, 6), idx),
        "sub tensor": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).get_sub_tensor([1, 2], [-1, 5]).get_local_tensor(),
        "local device": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).get_local_tensor(),
        "local host": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).get_local_tensor(host_view=True),
        "global device": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).get_global_tensor(),
        "global host": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).get_global_tensor(host_view=True),
        "from one file": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).from_filelist("a.bin"),
        "from files": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).from_filelist(["a", "b"], 0),
        "from prefix": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).from_file_prefix("pre"),
        "from prefix, 3 parts": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).from_file_prefix("pre", 3),
        "to file": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).local_to_file("x.bin"),
        "to prefix": lambda m: m.WholeMemoryTensor(_FakeWmbTensor(layers.log)).to_file_prefix("out"),
    }
    for name, script in scripts.items():
        result, log = _both(layers, script)
        if name in ("gather", "to prefix", "from prefix"):
            assert result[0] == "ok" and log, name
    # the comm of a tensor: both wrap the handle's communicator
    assert layers.ours.WholeMemoryTensor(_FakeWmbTensor([])).get_comm().get_size() == layers.ref.WholeMemoryTensor(_FakeWmbTensor([])).get_comm().get_size() == 5


def test_factories_make_the_same_binding_calls(layers, tmp_path):
    comm = types.SimpleNamespace(wmb_comm="the-comm")
    cases = [([10], torch.int64, None, None), ([10, 4], torch.float32, None, None), ([10, 4], torch.float16, [8, 1], [3, 3, 4]),
             ([10, 4], torch.float32, [3, 1], None), ([10, 4], torch.float32, [4, 2], None), ([2, 3, 4], torch.float32, None, None),
             ([], torch.float32, None, None), ([10], torch.int8, [1], None), ([10, 4], torch.bfloat16, [4], None)]
    for sizes, dtype, strides, partition in cases:
        for mt, loc in (("chunked", "cuda"), ("continuous", "cpu"), ("bogus", "cuda"), ("distributed", "moon")):
            _both(layers, lambda m: m.create_wholememory_tensor(comm, mt, loc, sizes, dtype, strides, partition).shape)
    a, b, odd = tmp_path / "a.bin", tmp_path / "b.bin", tmp_path / "odd.bin"
    a.write_bytes(b"\0" * (4 * 6 * 10))
    b.write_bytes(b"\0" * (4 * 6 * 3))
    odd.write_bytes(b"\0" * 50)
    for files, dtype, last, stride in ((str(a), torch.float32, 6, -1), ([str(a), str(b)], torch.float32, 6, 8), ([str(a), str(b)], torch.float32, 0, -1),
                                       ([str(a)], torch.int64, 3, -1), ([str(odd)], torch.float32, 6, -1), ([str(a), str(odd)], torch.float32, 0, -1),
                                       ([str(tmp_path / "missing")], torch.float32, 6, -1)):
        _both(layers, lambda m: m.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", files, dtype, last, stride).shape)
    _both(layers, lambda m: m.destroy_wholememory_tensor(m.WholeMemoryTensor(_FakeWmbTensor(layers.log))))
