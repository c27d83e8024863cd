"""multiprocess_run: one spawned Python process per rank, as the reference's tests and examples use it
(pylibwholegraph/utils/multiprocess.py:17-38).  Spawn, not fork: every rank creates its own CUDA context."""
import multiprocessing


def multiprocess_run(world_size: int, func, inline_single_process=False):
    """Run func(rank, world_size) in `world_size` processes and wait for them; a non-zero exit code of any rank is an
    AssertionError.  With world_size == 1 and inline_single_process the function runs in the calling process."""
    assert world_size > 0
    if inline_single_process and world_size == 1:
        func(0, 1)
        return
    ctx = multiprocessing.get_context("spawn")
    ranks = [ctx.Process(target=func, args=(rank, world_size)) for rank in range(world_size)]
    for p in ranks:
        p.start()
    for p in ranks:
        p.join()
    failed = [rank for rank, p in enumerate(ranks) if p.exitcode != 0]
    assert not failed, "ranks %s exited with an error" % failed
