#!/usr/bin/env python
"""Device-timed micro-benchmarks of the other hot-path kernels (--what scatter,adam,refadam,sample and, for the rows
DESIGN.md section 9 lists as not yet timed, unique,selfloop,weighted,budget) with their roofline arithmetic.  One GPU, world_size 1.  Not the headline metric (that is bench.py)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
import wholegraph_b200.binding as wmb  # noqa: E402
import wholegraph_b200.torch as wgth  # noqa: E402
from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="scatter,adam,sample")
    ap.add_argument("--rows", type=int, default=20_000_000)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
    comm = wgth.WholeMemoryCommunicator(wmb.create_communicator(wmb.create_unique_id(), 0, 1))
    env = get_wholegraph_env_fns()
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    out = []
    if "scatter" in args.what:
        rows, dim, n = args.rows, 256, 1 << 20
        t = wgth.create_wholememory_tensor(comm, "continuous", "cuda", [rows, dim], torch.float32, [dim, 1])
        src = torch.randn(n, dim, device="cuda")
        idxs = [torch.randperm(rows, device="cuda", generator=g)[:n].contiguous() for _ in range(4)]
        w_src = wrap_torch_tensor(src)
        w_idx = [wrap_torch_tensor(i) for i in idxs]
        k = [0]

        def f():
            wmb.wholememory_scatter_op(w_src, w_idx[k[0] % 4], t.wmb_tensor, env, get_stream())
            k[0] += 1
        ms = timeit(f)
        alg = n * (dim * 4 * 2 + 8)
        out.append({"op": "scatter fp32 %dx%d, %d rows" % (rows, dim, n), "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1), "frac_hbm": round(alg / ms / 1e6 / HBM, 4)})
        wgth.destroy_wholememory_tensor(t)
    if "adam" in args.what:
        rows, dim, n = args.rows // 4, 512, 1 << 18
        emb = wgth.create_embedding(comm, "distributed", "cuda", torch.float32, [rows, dim])
        opt = wgth.create_wholememory_optimizer(emb, "adam", {}, global_comm=comm)
        grads = torch.randn(n, dim, device="cuda")
        import numpy as np
        zipf = torch.from_numpy(((np.random.default_rng(5).zipf(1.05, size=n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) % np.uint64(rows)).astype(np.int64)).cuda()
        for name, idx in (("uniform", torch.randint(0, rows, (n,), device="cuda", generator=g)),
                          ("unique", torch.randperm(rows, device="cuda", generator=g)[:n].contiguous()),
                          ("zipf(1.05): hottest row gets %d gradients" % int(torch.bincount(zipf).max().item()), zipf)):
            w_i, w_g = wrap_torch_tensor(idx), wrap_torch_tensor(grads)
            uniq = int(torch.unique(idx).numel())

            def f():
                wmb.EmbeddingGatherGradientApply(emb.wmb_embedding, w_i, w_g, False, 0.01, env, get_stream())
            ms = timeit(f, steps=10, warmup=3)
            alg = n * dim * 4 + uniq * (6 * dim * 4 + 16) + n * 8
            out.append({"op": "fused dedup+LazyAdam %dx%d, %d grads (%s, %d unique), whole call incl. sort" % (rows, dim, n, name, uniq),
                        "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1), "frac_hbm": round(alg / ms / 1e6 / HBM, 4)})
        wgth.destroy_wholememory_optimizer(opt)
        wgth.destroy_embedding(emb)
    if "refadam" in args.what:
        # The REFERENCE's own dedup + LazyAdam kernels (oracle/_ref build, oracle/ref_optimizer_hook.cpp) on the workload of the
        # "adam" entry above.  Run with WHOLEGRAPH_B200_LIB=oracle/_ref/libwholegraph_ref.so python tools/bench_ops.py --what refadam
        # The hook cudaMallocs its two dedup buffers and synchronises the stream per call (the reference pipeline also
        # synchronises, embedding.cpp:146-323); both are inside the timed region, as the sort is inside ours.
        import ctypes
        from wholegraph_b200 import _lib
        fn = _lib.lib.wgref_dedup_and_optimizer_step
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int64] + [ctypes.c_float] * 4 + [ctypes.c_int] + \
                      [ctypes.c_float] * 2 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        rows, dim, n = args.rows // 4, 512, 1 << 18
        w = torch.randn(rows, dim, device="cuda")
        state = torch.zeros(rows, 2 * dim, device="cuda")
        b12 = torch.ones(rows, 2, device="cuda")
        grads = torch.randn(n, dim, device="cuda")
        w_w, w_s, w_b, w_g = (wrap_torch_tensor(t) for t in (w, state, b12, grads))
        for name, idx in (("uniform", torch.randint(0, rows, (n,), device="cuda", generator=g)),
                          ("unique", torch.randperm(rows, device="cuda", generator=g)[:n].contiguous())):
            w_i = wrap_torch_tensor(idx)
            uniq = int(torch.unique(idx).numel())
            dd = ctypes.c_int64(0)

            def f():
                rc = fn(2, w_i.get_c_handle(), w_g.get_c_handle(), w_w.get_c_handle(), w_s.get_c_handle(), w_b.get_c_handle(), 0,
                        0.0, 1e-8, 0.9, 0.999, 0, 0.99, 0.01, env, get_stream(), ctypes.byref(dd))
                assert rc == 0
            ms = timeit(f, steps=10, warmup=3)
            assert dd.value == uniq
            alg = n * dim * 4 + uniq * (6 * dim * 4 + 16) + n * 8
            out.append({"op": "REFERENCE dedup + LazyAdam kernels %dx%d, %d grads (%s, %d unique), hook call" % (rows, dim, n, name, uniq),
                        "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1), "frac_hbm": round(alg / ms / 1e6 / HBM, 4)})
    if "sample" in args.what:
        nodes = 10_000_000
        deg = torch.clamp((torch.rand(nodes, device="cuda", generator=g) ** -0.7).long(), max=10000)
        row_ptr = torch.zeros(nodes + 1, dtype=torch.int64, device="cuda")
        row_ptr[1:] = torch.cumsum(deg, 0)
        edges = int(row_ptr[-1].item())
        rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
        cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [edges], torch.int32, [1])
        rp.get_local_tensor()[0].copy_(row_ptr)
        cp.get_local_tensor()[0].copy_(torch.randint(0, nodes, (edges,), device="cuda", dtype=torch.int32, generator=g))
        for ncenter, k in ((1024, 25), (25 * 1024, 10), (262144, 25), (262144, 10)):
            centers = torch.randint(0, nodes, (ncenter,), device="cuda", generator=g)

            def f():
                return wgth.unweighted_sample_without_replacement(rp.wmb_tensor, cp.wmb_tensor, centers, k, random_seed=7)
            ms = timeit(f, steps=10, warmup=3)
            total = int(f()[0][-1].item())
            out.append({"op": "unweighted sample %d centers k=%d on %d-node/%d-edge CSR (whole call: count+scan+sync+sample)" % (ncenter, k, nodes, edges),
                        "ms": round(ms, 4), "Msamples_per_s": round(total / ms / 1e3, 2), "samples": total})
        wgth.destroy_wholememory_tensor(rp)
        wgth.destroy_wholememory_tensor(cp)
    if "unique" in args.what:
        # f1: the op between two sampling hops -- targets = one hop's centers, neighbors = its samples (C5 shape: 25 per center)
        for targets, per in ((1024, 25), (25 * 1024, 10), (262144, 25)):
            tg = torch.randperm(50_000_000, device="cuda", generator=g)[:targets].contiguous()
            nb = torch.randint(0, 50_000_000, (targets * per,), device="cuda", generator=g)

            def f():
                return wgth.graph_ops.append_unique(tg, nb, need_neighbor_raw_to_unique=True)
            ms = timeit(f, steps=10, warmup=3)
            uniq = int(f()[0].shape[0])
            alg = (targets + targets * per) * 8 + uniq * 8 + targets * per * 4
            out.append({"op": "append_unique %d targets + %d neighbors int64 -> %d unique (+ mapping), whole call" % (targets, targets * per, uniq),
                        "ms": round(ms, 4), "Mids_per_s": round((targets + targets * per) / ms / 1e3, 2), "alg_GBps": round(alg / ms / 1e6, 1)})
    if "selfloop" in args.what:
        rows = 262144
        deg = torch.randint(0, 26, (rows,), device="cuda", generator=g)
        rp32 = torch.zeros(rows + 1, dtype=torch.int32, device="cuda")
        rp32[1:] = torch.cumsum(deg, 0).to(torch.int32)
        col32 = torch.randint(0, rows * 4, (int(rp32[-1].item()),), device="cuda", dtype=torch.int32, generator=g)

        def f():
            return wgth.graph_ops.add_csr_self_loop(rp32, col32)
        ms = timeit(f, steps=10, warmup=3)
        alg = 2 * (rows + 1) * 4 + col32.numel() * 4 * 2 + rows * 4
        out.append({"op": "csr_add_self_loop %d rows, %d edges" % (rows, col32.numel()), "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1)})
    if "weighted" in args.what:
        nodes = 10_000_000
        deg = torch.clamp((torch.rand(nodes, device="cuda", generator=g) ** -0.7).long(), max=10000)
        row_ptr = torch.zeros(nodes + 1, dtype=torch.int64, device="cuda")
        row_ptr[1:] = torch.cumsum(deg, 0)
        edges = int(row_ptr[-1].item())
        rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
        cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [edges], torch.int32, [1])
        wp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [edges], torch.float32, [1])
        rp.get_local_tensor()[0].copy_(row_ptr)
        cp.get_local_tensor()[0].copy_(torch.randint(0, nodes, (edges,), device="cuda", dtype=torch.int32, generator=g))
        wp.get_local_tensor()[0].copy_(torch.rand(edges, device="cuda", generator=g) + 0.01)
        for ncenter, k in ((1024, 25), (262144, 25), (262144, 10)):
            centers = torch.randint(0, nodes, (ncenter,), device="cuda", generator=g)

            def f():
                return wgth.weighted_sample_without_replacement(rp.wmb_tensor, cp.wmb_tensor, wp.wmb_tensor, centers, k, random_seed=7)
            ms = timeit(f, steps=10, warmup=3)
            total = int(f()[0][-1].item())
            out.append({"op": "weighted sample %d centers k=%d on %d-node/%d-edge CSR (whole call)" % (ncenter, k, nodes, edges),
                        "ms": round(ms, 4), "Msamples_per_s": round(total / ms / 1e3, 2), "samples": total})
        for t in (rp, cp, wp):
            wgth.destroy_wholememory_tensor(t)
    if "budget" in args.what:
        # f2: gather under an SM budget (the persistent-grid mode a loader uses to overlap sampling with the gather)
        rows, dim, n = args.rows, 256, 1 << 20
        t = wgth.create_wholememory_tensor(comm, "continuous", "cuda", [rows, dim], torch.float32, [dim, 1])
        idxs = [torch.randint(0, rows, (n,), device="cuda", generator=g) for _ in range(4)]
        dst = torch.empty(n, dim, device="cuda")
        w_dst, w_idx = wrap_torch_tensor(dst), [wrap_torch_tensor(i) for i in idxs]
        for sms in (-1, 132, 96, 64, 32):
            k = [0]

            def f():
                wmb.wholememory_gather_op(t.wmb_tensor, w_idx[k[0] % 4], w_dst, env, get_stream(), sms)
                k[0] += 1
            ms = timeit(f)
            alg = n * (dim * 4 * 2 + 8)
            out.append({"op": "gather fp32 %dx%d, %d rows, gather_sms=%d" % (rows, dim, n, sms), "ms": round(ms, 4),
                        "alg_GBps": round(alg / ms / 1e6, 1), "frac_hbm": round(alg / ms / 1e6 / HBM, 4)})
        wgth.destroy_wholememory_tensor(t)
    if "c1" in args.what:
        # a7 / BASELINE config C1: wholememory_gather 1M x 64 fp32, single rank, HOST-location memory, 100,000 int64 ids.
        # The reference sorts the ids first for host memory (gather_op.cpp:116-120, sort_indices_func.cu:42-92); this library
        # does not.  Timed on whichever library is loaded, random ids and the same ids pre-sorted by the caller (= what a sort
        # inside the call could buy at best, without its cost).
        rows, dim, n = 1_000_000, 64, 100_000
        t = wgth.create_wholememory_tensor(comm, "continuous", "cpu", [rows, dim], torch.float32, [dim, 1])
        idx = torch.randint(0, rows, (n,), device="cuda", generator=g)
        dst = torch.empty(n, dim, device="cuda")
        w_dst = wrap_torch_tensor(dst)
        for name, ids in (("random ids", idx), ("caller-sorted ids", torch.sort(idx)[0].contiguous())):
            w_i = wrap_torch_tensor(ids)

            def f():
                wmb.wholememory_gather_op(t.wmb_tensor, w_i, w_dst, env, get_stream())
            ms = timeit(f, steps=30, warmup=5)
            out.append({"op": "C1 gather HOST/CONTINUOUS %dx%d fp32, %d int64 ids, %s" % (rows, dim, n, name), "ms": round(ms, 4),
                        "GBps_out": round(n * dim * 4 / ms / 1e6, 2)})
        wgth.destroy_wholememory_tensor(t)
    if "overlap" in args.what:
        # f2: a loader overlaps feature gathering with the next batch's sampling by giving the gather an SM budget
        # (gather_sms; reference gather_scatter_func.cuh:440,506).  Gather on one stream, unweighted sampling on another:
        # alone, back to back, and concurrent for several budgets.
        rows, dim, n = args.rows, 256, 1 << 20
        t = wgth.create_wholememory_tensor(comm, "continuous", "cuda", [rows, dim], torch.float32, [dim, 1])
        idxs = [torch.randint(0, rows, (n,), device="cuda", generator=g) for _ in range(4)]
        dst = torch.empty(n, dim, device="cuda")
        w_dst, w_idx = wrap_torch_tensor(dst), [wrap_torch_tensor(i) for i in idxs]
        nodes = 10_000_000
        deg = torch.clamp((torch.rand(nodes, device="cuda", generator=g) ** -0.7).long(), max=10000)
        row_ptr = torch.zeros(nodes + 1, dtype=torch.int64, device="cuda")
        row_ptr[1:] = torch.cumsum(deg, 0)
        edges = int(row_ptr[-1].item())
        rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
        cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [edges], torch.int32, [1])
        rp.get_local_tensor()[0].copy_(row_ptr)
        cp.get_local_tensor()[0].copy_(torch.randint(0, nodes, (edges,), device="cuda", dtype=torch.int32, generator=g))
        centers = torch.randint(0, nodes, (262144,), device="cuda", generator=g)
        s_gather, s_sample = torch.cuda.Stream(), torch.cuda.Stream()
        reps = 8  # gathers / sampling calls per measurement

        def gathers(sms):
            with torch.cuda.stream(s_gather):
                for i in range(reps):
                    wmb.wholememory_gather_op(t.wmb_tensor, w_idx[i % 4], w_dst, env, get_stream(), sms)

        def samples():
            with torch.cuda.stream(s_sample):
                for _ in range(reps):
                    wgth.unweighted_sample_without_replacement(rp.wmb_tensor, cp.wmb_tensor, centers, 25, random_seed=7)

        def wall(fn):
            import time
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / 3 * 1e3

        t_g = wall(lambda: gathers(-1))
        t_s = wall(samples)
        out.append({"op": "overlap: %d gathers (fp32 %dx%d, %d rows) alone, all SMs" % (reps, rows, dim, n), "ms": round(t_g, 3)})
        out.append({"op": "overlap: %d sampling calls (262144 centers, k=25) alone" % reps, "ms": round(t_s, 3)})
        for sms in (-1, 132, 116, 96, 64):
            t_both = wall(lambda: (gathers(sms), samples()))
            out.append({"op": "overlap: both, two streams, gather_sms=%d" % sms, "ms": round(t_both, 3), "serial_sum_ms": round(t_g + t_s, 3),
                        "vs_serial": round((t_g + t_s) / t_both, 3)})
        for x in (t, rp, cp):
            wgth.destroy_wholememory_tensor(x)
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
