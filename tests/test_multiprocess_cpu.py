"""world_size-2 tests of the multi-rank HOST logic on a CPU box: torch.distributed (gloo) carries the
unique id exactly like the reference's comm.py, the library's own AF_UNIX bootstrap does the rest
(rank discovery, collectives, fd passing, split).  No GPU needed."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import wholegraph_b200.binding as wmb
    import wholegraph_b200.torch as wgth
    from wholegraph_b200 import _lib
    try:
        wgth.init_torch_env(rank, world_size, rank, world_size, wm_log_level="warn", backend="gloo")
        comm = wgth.get_global_communicator()
        assert comm.get_rank() == rank and comm.get_size() == world_size
        assert wgth.get_local_node_communicator() is comm
        # every rank must agree that nothing device-mapped can be allocated on a GPU-less box
        assert comm.support_type_location("distributed", "cuda")
        assert comm.support_type_location("continuous", "cuda") == torch.cuda.is_available()
        # split: ranks with the same color form a new communicator ordered by key
        sub = wgth.split_communicator(comm, color=rank % 2, key=0)
        assert sub.get_size() == (world_size + 1 - rank % 2) // 2 and sub.get_rank() == rank // 2
        sub.destroy()
        # key decides the order inside the new communicator (ties by old rank): descending keys reverse the ranks
        rev = wgth.split_communicator(comm, color=0, key=world_size - rank)
        assert rev.get_size() == world_size and rev.get_rank() == world_size - 1 - rank
        rev.barrier()
        rev.destroy()
        # a negative color opts out (the Python helper answers None without entering the collective, like the reference's)
        assert wgth.split_communicator(comm, color=-1) is None
        # a second communicator in the reverse order of creation still matches ranks up
        comm2 = wgth.create_group_communicator()
        assert comm2.get_rank() == rank
        wgth.destroy_communicator(comm2)
        # strided groups (reference comm.py:133): 4 ranks, group_size 2, stride 2 -> [0, 2] and [1, 3]
        if world_size == 4:
            g = wgth.create_group_communicator(2, 2)
            assert g.get_size() == 2 and g.get_rank() == rank // 2
            g.barrier()
            wgth.destroy_communicator(g)
            g = wgth.create_group_communicator(2, 1)  # [0, 1] and [2, 3]
            assert g.get_size() == 2 and g.get_rank() == rank % 2
            g.barrier()
            wgth.destroy_communicator(g)
            dev = wgth.get_local_device_communicator()
            assert dev.get_size() == 1 and dev.get_rank() == 0 and wgth.get_local_device_communicator() is dev
        # collective argument check: different sizes on different ranks must be rejected, not deadlock
        if not torch.cuda.is_available():
            try:
                wmb.malloc(1024 * (rank + 1), comm.wmb_comm, wmb.MtDistributed, wmb.MlDevice, 64)
                ok = False
            except RuntimeError:
                ok = True
            assert ok
            # same arguments everywhere, but no rank CAN allocate (no GPU): every rank gets the error -- and stays in step:
            # the failure travels through the collectives of the allocation (fail_together) instead of skipping them
            for mt, ml in ((wmb.MtChunked, wmb.MlDevice), (wmb.MtContinuous, wmb.MlHost), (wmb.MtDistributed, wmb.MlHost)):
                try:
                    wmb.malloc(1 << 20, comm.wmb_comm, mt, ml, 64)
                    ok = False
                except (RuntimeError, NotImplementedError) as e:  # CUDA error, or NOT_SUPPORTED (code 9 -> "not recognized")
                    ok = "CUDA" in str(e) or "not recognized" in str(e)
                assert ok, (mt, ml)
                comm.barrier()  # still in step after the failed collective
            # different custom partitions with the same total: LOGIC_ERROR on every rank (a hash of the array is compared)
            part = [8] * world_size
            if rank == 1:
                part[0], part[-1] = 7, 9
            try:
                wmb.malloc(8 * world_size * 64, comm.wmb_comm, wmb.MtDistributed, wmb.MlDevice, 64, part)
                ok = False
            except RuntimeError as e:
                ok = "ogic" in str(e)
            assert ok
            comm.barrier()
        results[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        results[rank] = "FAIL: " + traceback.format_exc()
    finally:
        try:
            wgth.finalize()
        except Exception:
            pass


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world_size", [2, 3, 4])
def test_control_plane_world_size_n_over_gloo(world_size):
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    results = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, results)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("rank hung")
    assert dict(results) == {r: "ok" for r in range(world_size)}, dict(results)


def _timeout_worker(rank, world_size, port, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WG_BOOTSTRAP_TIMEOUT_S"] = "2"
    import time
    import wholegraph_b200.torch as wgth
    try:
        wgth.init_torch_env(rank, world_size, rank, world_size, wm_log_level="warn", backend="gloo")
        comm = wgth.get_global_communicator()
        comm.barrier()  # everybody: fine
        if rank == 0:
            t0 = time.time()
            try:
                comm.barrier()  # rank 1 never joins this one
                results[rank] = "FAIL: barrier returned"
            except RuntimeError as e:
                dt = time.time() - t0
                results[rank] = "ok" if 1.0 < dt < 30.0 else "FAIL: gave up after %.1f s (%s)" % (dt, e)
        else:
            time.sleep(8)  # alive, socket open, but on a different code path
            results[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        results[rank] = "FAIL: " + traceback.format_exc()


@pytest.mark.timeout(120)
def test_collective_mismatch_ends_in_an_error_not_a_hang():
    """A rank that waits for a collective its peer never issues must get WHOLEMEMORY_COMMUNICATION_ERROR after
    WG_BOOTSTRAP_TIMEOUT_S, not block forever (a hung GPU box costs a whole run)."""
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    results = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("rank hung")
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)
