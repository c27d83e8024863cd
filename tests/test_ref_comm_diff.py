"""pylibwholegraph/torch/comm.py of the reference (loaded unchanged through compat/) next to wholegraph_b200/torch/comm.py,
every rank of a job simulated in one process over fakes of torch.distributed and of the binding's communicator calls:

 * create_group_communicator(group_size, comm_stride): for every rank of worlds 1..24 and every valid (group_size, stride),
   both implementations issue the same sequence of id broadcasts (same roots, so the two could even be mixed in one job)
   and join the same communicator: the same root's unique id, the same rank inside the group, the same size;
 * get_global / get_local_node / get_local_device communicator: which of them are one and the same object, for every world
   / node shape and every order of first use.

CPU only."""
import itertools
import os

import pytest
import torch

import wholegraph_b200.binding as wmb
from wholegraph_b200.torch import comm as our_mod

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/comm.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


class _FakeDist:
    def __init__(self, log):
        self.rank, self.world, self.log = 0, 1, log

    def is_initialized(self):
        return True

    def get_rank(self):
        return self.rank

    def get_world_size(self):
        return self.world

    def get_backend(self):
        return "gloo"

    def broadcast(self, tensor, src):
        self.log.append(("broadcast", src))
        tensor.fill_(src + 1)  # what the root's id looks like to everybody: a function of the root alone


class _FakeComm:
    def __init__(self, root, rank, size):
        self.root, self.rank, self.size, self.backend = root, rank, size, None


@pytest.fixture()
def job(monkeypatch):
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF, "_reference_comm", package="pylibwholegraph.torch")
    log = []
    dist = _FakeDist(log)
    real_uid = wmb.create_unique_id

    def fake_create_unique_id():
        uid = real_uid()
        uid.as_tensor().fill_(dist.rank + 1)
        return uid

    def fake_create_communicator(uid, rank, size):
        root = int(uid.as_tensor()[0]) - 1
        assert bool((uid.as_tensor() == root + 1).all())
        log.append(("join", root, rank, size))
        return _FakeComm(root, rank, size)

    def fake_set_backend(comm, backend):
        comm.backend = backend

    monkeypatch.setattr(wmb, "create_unique_id", fake_create_unique_id)
    monkeypatch.setattr(wmb, "create_communicator", fake_create_communicator)
    monkeypatch.setattr(wmb, "communicator_set_distributed_backend", fake_set_backend)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self.clone())  # the reference stages ids on the GPU
    for mod in (our_mod, ref_mod):
        monkeypatch.setattr(mod, "dist", dist)
        mod.reset_communicators()
    yield types_ns(ours=our_mod, ref=ref_mod, dist=dist, log=log)
    for mod in (our_mod, ref_mod):
        mod.reset_communicators()


def types_ns(**kw):
    import types
    return types.SimpleNamespace(**kw)


def test_group_communicators_every_rank_every_shape(job):
    checked = 0
    for world in (1, 2, 3, 4, 6, 8, 12, 16, 24):
        job.dist.world = world
        shapes = [(-1, 1)] + [(g, s) for g in range(1, world + 1) for s in range(1, world + 1) if world % (g * s) == 0]
        for group_size, stride in shapes:
            for rank in range(world):
                job.dist.rank = rank
                traces = []
                for mod in (job.ours, job.ref):
                    del job.log[:]
                    comm = mod.create_group_communicator(group_size, stride)
                    # (a one-rank job needs no exchange: this repo skips the self-broadcast, which also lets it run without a
                    # process group; the reference issues it anyway)
                    calls = [c for c in job.log if not (world == 1 and c[0] == "broadcast")]
                    traces.append((calls, comm.wmb_comm.root, comm.wmb_comm.rank, comm.wmb_comm.size))
                assert traces[0] == traces[1], (world, group_size, stride, rank, traces)
                g = world if group_size == -1 else group_size
                _log, root, r, size = traces[0]
                assert size == g and root + r * stride == rank  # member r of a group sits r strides after its root
                checked += 1
    assert checked > 1400


@pytest.mark.parametrize("world,local", [(1, 1), (8, 8), (16, 8), (4, 1), (2, 2)])
def test_well_known_communicators_share_objects_like_the_reference(job, world, local):
    getters = ("get_global_communicator", "get_local_node_communicator", "get_local_device_communicator")
    for order in itertools.permutations(range(3)):
        shared = []
        for mod in (job.ours, job.ref):
            mod.reset_communicators()
            job.dist.world, job.dist.rank = world, world - 1
            mod.set_world_info(world - 1, world, (world - 1) % local, local)
            first = {}
            for i in order:
                first[i] = getattr(mod, getters[i])()
            again = {i: getattr(mod, getters[i])() for i in range(3)}
            assert all(first[i] is again[i] for i in range(3))                      # cached
            shared.append(tuple(again[a] is again[b] for a, b in ((0, 1), (0, 2), (1, 2))))
            sizes = tuple(again[i].wmb_comm.size for i in range(3))
            assert sizes == (world, local, 1)
            shared.append(tuple(again[i].wmb_comm.backend for i in range(3)))   # which of them had the backend set explicitly
        assert shared[:2] == shared[2:], (world, local, order, shared)
