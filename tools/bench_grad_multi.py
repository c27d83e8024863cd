#!/usr/bin/env python
"""Multi-GPU gradient-apply micro-benchmark (config C4's write path): every rank applies `--grads` fp32 gradient rows of
width `--dim` with uniform global ids to an N x `--rows-per-gpu` DISTRIBUTED LazyAdam embedding.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             tools/bench_grad_multi.py
WG_GRAD_PUSH=0 selects the NCCL all-to-all exchange instead of the peer-store push.  Device-timed, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
import wholegraph_b200.binding as wmb  # noqa: E402
import wholegraph_b200.torch as wgth  # noqa: E402
from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=5_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--grads", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--zipf", type=float, default=0.0, help="> 1: ids drawn Zipf(a) over the rows (hot rows, duplicates) instead of uniform")
    args = ap.parse_args()
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
    uid = wmb.create_unique_id() if rank == 0 else wmb.PyWholeMemoryUniqueID()
    if world > 1:
        t = uid.as_tensor().cuda()
        dist.broadcast(t, 0)
        uid.as_tensor().copy_(t.cpu())
    comm = wgth.WholeMemoryCommunicator(wmb.create_communicator(uid, rank, world))
    rows = args.rows_per_gpu * world
    emb = wgth.create_embedding(comm, "distributed", "cuda", torch.float32, [rows, args.dim])
    opt = wgth.create_wholememory_optimizer(emb, "adam", {}, global_comm=comm)
    g = torch.Generator(device="cuda")
    g.manual_seed(100 + rank)
    grads = torch.randn(args.grads, args.dim, device="cuda", generator=g)
    if args.zipf > 1.0:
        import numpy as np
        nrng = np.random.default_rng(100 + rank)
        # hot rows spread over the whole table (a fixed odd multiplier scatters the small Zipf values across all owners)
        idxs = [torch.from_numpy(((nrng.zipf(args.zipf, size=args.grads).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) % np.uint64(rows)).astype(np.int64)).cuda()
                for _ in range(4)]
    else:
        idxs = [torch.randint(0, rows, (args.grads,), device="cuda", generator=g) for _ in range(4)]
    w_g = wrap_torch_tensor(grads)
    w_i = [wrap_torch_tensor(i) for i in idxs]
    env = get_wholegraph_env_fns()

    def step(k):
        wmb.EmbeddingGatherGradientApply(emb.wmb_embedding, w_i[k % 4], w_g, False, 0.01, env, get_stream())

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        row_bytes = args.dim * 4
        print(json.dumps({"op": "gradient apply (LazyAdam)", "n_gpus": world, "push": os.environ.get("WG_GRAD_PUSH", "1") != "0",
                          "grads_per_rank": args.grads, "dim": args.dim, "rows_total": rows, "ids": "zipf(%g)" % args.zipf if args.zipf > 1.0 else "uniform",
                          "unique_ids_rank0_batch0": int(torch.unique(idxs[0]).numel()), "ms_per_step": round(ms, 4),
                          "Mrows_per_s_total": round(args.grads * world / ms / 1e3, 2),
                          "grad_GBps_per_gpu": round(args.grads * row_bytes / ms / 1e6, 1)}))
    wgth.destroy_wholememory_optimizer(opt)
    wgth.destroy_embedding(emb)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
