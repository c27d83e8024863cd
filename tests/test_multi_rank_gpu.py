"""Multi-rank tests on the GPU box.  world_size 1 always runs; world_size 2/4 run with one rank per GPU when the
box has enough GPUs, and with ranks SHARING GPU 0 for the mapped-memory scenarios (no NCCL involved) otherwise."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, scenario, env=None, share_gpu=False):
    import multi_rank_scenarios as S
    ngpus = torch.cuda.device_count()
    if share_gpu:
        ngpus = 1
    elif ngpus < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = _free_port()
    # a rank that diverges must end in an error, not park its peers in a collective until the join below gives up
    env = dict({"WG_BOOTSTRAP_TIMEOUT_S": "300"}, **(env or {}))
    procs = [ctx.Process(target=S.worker, args=(r, world, port, ngpus, scenario, env, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    assert not hung, "rank hung"
    assert dict(results) == {r: "ok" for r in range(world)}, "\n".join(f"[{k}] {v}" for k, v in dict(results).items())


@pytest.mark.timeout(900)
@pytest.mark.parametrize("scenario", ["gather_scatter", "gradient", "sampling", "file_io", "weighted_sampling"])
def test_single_rank(scenario):
    _run(1, scenario)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("scenario", ["gather_scatter", "sampling"])
def test_single_rank_forced_bucket_exchange(scenario):
    """DISTRIBUTED ops through the partition + exchange path (self-exchange on one rank)."""
    _run(1, scenario, env={"WG_FORCE_EXCHANGE": "1"})


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("scenario", ["gather_scatter", "gradient", "sampling", "file_io", "weighted_sampling"])
def test_ranks_sharing_one_gpu_mapped_memory(world, scenario):
    """Cross-process VMM mapping (POSIX fd over AF_UNIX), partitions, peer addressing, peer-store gradient push --
    no NCCL needed."""
    _run(world, scenario, share_gpu=True)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_rank_local_failures_reach_every_rank(world):
    """A missing part file on one rank / different partitions on different ranks: every rank gets an error, none hangs."""
    _run(world, "failing_together", share_gpu=world > 1)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("scenario", ["gather_scatter", "gradient", "sampling", "file_io"])
def test_one_rank_per_gpu(world, scenario):
    _run(world, scenario)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("scenario", ["gather_scatter", "sampling"])
def test_one_rank_per_gpu_forced_nccl_exchange(world, scenario):
    _run(world, scenario, env={"WG_FORCE_EXCHANGE": "1"})


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_one_rank_per_gpu_gradient_over_nccl(world):
    """The all-to-all gradient exchange (used when GPUs cannot map each other) instead of the peer-store push."""
    _run(world, "gradient", env={"WG_GRAD_PUSH": "0"})


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2])
def test_one_rank_per_gpu_distributed_without_peer_mapping(world):
    _run(world, "gather_scatter", env={"WG_DISTRIBUTED_NO_PEER": "1"})
