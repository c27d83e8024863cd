#!/usr/bin/env python
"""Multi-GPU neighbor-sampling benchmark (BASELINE config C5): unweighted sample fanout=[25,10] on a synthetic CSR
held in CHUNKED/DEVICE WholeMemory sharded over the N GPUs of the box, `--seeds` seed nodes per rank per step.

Graph: `--nodes` nodes, degrees from a power law clipped to [0, 10000] and scaled towards `--edges`
edges in total, uniform int32 column ids, int64 row_ptr.  Every rank derives the same degree sequence from the same seed
and writes only its own shard of row_ptr / col_idx.
One step = GraphStructure.multilayer_sample_without_replacement(seeds, [25, 10]): hop 1 (k=25) -> append_unique ->
hop 2 (k=10) -> append_unique, all inside the library.  Device-timed (CUDA events), max over ranks.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             tools/bench_sample_multi.py [--nodes 111059956 --edges 1000000000]
Defaults are a 1/10-scale graph so a first run is cheap; pass the full sizes for the C5 figure."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=11_105_995)
    ap.add_argument("--edges", type=int, default=100_000_000)
    ap.add_argument("--seeds", type=int, default=1024)
    ap.add_argument("--fanout", default="25,10")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    fanout = [int(x) for x in args.fanout.split(",")]
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

    # identical degree sequence on every rank (same generator seed), built in slices to bound temporary memory
    nodes = args.nodes
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xC5)
    deg = torch.empty(nodes, dtype=torch.int64, device="cuda")
    chunk = 16_000_000
    for s in range(0, nodes, chunk):
        e = min(nodes, s + chunk)
        u = torch.rand(e - s, device="cuda", generator=gen).clamp_(min=1e-9)
        deg[s:e] = torch.clamp((u ** -0.7), max=10000.0).long()
    scale = args.edges / float(deg.sum().item())
    deg = torch.clamp((deg.double() * scale).round().long(), min=0, max=10000)
    row_ptr = torch.zeros(nodes + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(deg, 0, out=row_ptr[1:])
    edges = int(row_ptr[-1].item())
    del deg

    rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
    cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [edges], torch.int32, [1])
    local_rp, rp_first = rp.get_local_tensor()
    local_rp.copy_(row_ptr[rp_first:rp_first + local_rp.shape[0]])
    del row_ptr
    local_cp, _ = cp.get_local_tensor()
    gen.manual_seed(0xC500 + rank)
    for s in range(0, local_cp.shape[0], 64_000_000):
        e = min(local_cp.shape[0], s + 64_000_000)
        local_cp[s:e] = torch.randint(0, nodes, (e - s,), device="cuda", dtype=torch.int32, generator=gen)
    torch.cuda.synchronize()
    comm.barrier()

    graph = wgth.GraphStructure()
    graph.set_csr_graph(rp, cp)
    seeds = [torch.randint(0, nodes, (args.seeds,), device="cuda", dtype=torch.int32, generator=gen) for _ in range(4)]  # ids share the col dtype (append_unique needs one id type)

    def step(k):
        return graph.multilayer_sample_without_replacement(seeds[k % 4], fanout, random_seed=1234 + k)

    sampled = 0
    for k in range(max(args.warmup, 3)):
        target_gids, edge_indice, csr_row_ptr, csr_col_ind = step(k)
    sampled = sum(int(c.shape[0]) for c in csr_col_ind)
    frontier = [int(t.shape[0]) for t in target_gids]
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
        print(json.dumps({"op": "unweighted multi-hop sample fanout=%s" % fanout, "n_gpus": world, "nodes": nodes, "edges": edges,
                          "seeds_per_rank": args.seeds, "frontier_sizes_last_step_rank0": frontier,
                          "sampled_edges_per_step_rank0": sampled, "ms_per_step": round(ms, 4),
                          "Msamples_per_s_total": round(sampled * world / ms / 1e3, 2),
                          "native_env": os.environ.get("WG_TORCH_NATIVE_ENV", "0") == "1"}))
    wgth.destroy_wholememory_tensor(rp)
    wgth.destroy_wholememory_tensor(cp)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
