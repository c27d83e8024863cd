"""Multi-rank GPU scenarios run by test_multi_rank_gpu.py (one spawned process per rank).

Every rank rebuilds the SAME host-side model of the whole table from a shared seed, performs its part through the
C ABI, and checks what it can see against the oracle.  Ranks may share one GPU (VMM mappings work across
processes on the same device); scenarios that need NCCL require one GPU per rank."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(rank, world, port, ngpus, env):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(env)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["WG_TEST_SHARED_GPU"] = "1" if ngpus < world else "0"
    import torch
    import wholegraph_b200.torch as wgth
    torch.cuda.set_device(rank % ngpus)
    # gloo process group: only used to broadcast the unique id (ranks may share a GPU, which NCCL refuses)
    wgth.init_torch_env(rank, world, rank % ngpus, world, wm_log_level="warn", backend="gloo")
    torch.cuda.set_device(rank % ngpus)
    return wgth.get_global_communicator()


def _random_partition(rng, rows, world):
    cuts = np.sort(rng.choice(np.arange(1, rows), size=world - 1, replace=False))
    return np.diff(np.concatenate([[0], cuts, [rows]])).astype(np.int64).tolist()


def scenario_gather_scatter(rank, world, comm):
    import torch
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from oracle import oracle as O
    dev = torch.cuda.current_device()
    cases = [("continuous", "cuda", O.DT_FLOAT, O.DT_FLOAT, 128, 128, False), ("chunked", "cuda", O.DT_HALF, O.DT_HALF, 128, 128, False),
             ("chunked", "cuda", O.DT_FLOAT, O.DT_HALF, 33, 36, True), ("distributed", "cuda", O.DT_FLOAT, O.DT_FLOAT, 64, 64, False),
             ("distributed", "cuda", O.DT_INT64, O.DT_INT, 7, 8, True), ("continuous", "cpu", O.DT_FLOAT, O.DT_FLOAT, 32, 32, False),
             ("chunked", "cpu", O.DT_HALF, O.DT_FLOAT, 11, 12, True), ("continuous", "cuda", O.DT_DOUBLE, O.DT_DOUBLE, 5, 5, True)]
    rows = 10007
    for ci, (mt, loc, tab_dt, out_dt, cols, stride, custom) in enumerate(cases):
        rng = np.random.default_rng(1000 + ci)  # same stream on every rank
        part = _random_partition(rng, rows, world) if custom else None
        host = G.random_table(rng, tab_dt, rows, stride)
        t = wmb.create_wholememory_matrix(G.WM_OF[tab_dt], rows, cols, stride, comm.wmb_comm, G.MT[mt], G.ML[loc], part)
        view_loc = wmb.MlDevice if loc == "cuda" else wmb.MlHost
        local, first = t.get_wholememory_handle().get_local_flatten_tensor(G.WM_OF[tab_dt], view_loc, dev)
        first_row = first // stride
        nloc = local.numel() // stride
        offs = t.get_entry_offsets()
        assert offs[rank] == first_row and offs[rank + 1] - offs[rank] == nloc and offs[-1] == rows, (offs, first_row, nloc)
        if part is not None:
            assert np.diff(offs).tolist() == part
        local.reshape(nloc, stride).copy_(G.np_to_torch(host[first_row:first_row + nloc], tab_dt))
        comm.barrier()
        rrng = np.random.default_rng(77 * ci + rank)
        idx = rrng.integers(0, rows, size=3000 + rank).astype(np.int64 if ci % 2 == 0 else np.int32)
        idx[::131] = -1
        n = idx.shape[0]
        sentinel = G.random_table(rrng, out_dt, n, cols)
        out_t = G.np_to_torch(sentinel.copy(), out_dt).cuda()
        G.gather(t, torch.from_numpy(idx).cuda(), out_t)
        torch.cuda.synchronize()
        exp = sentinel.copy()
        O.gather(host, tab_dt, idx, out_dt, out=exp, cols=cols)
        assert G.torch_to_np(out_t, out_dt).tobytes() == exp.tobytes(), f"gather mismatch case {ci} rank {rank}"
        comm.barrier()
        # scatter: rank r owns target rows r, r+world, ... (disjoint across ranks), input in the OUT dtype family
        target = np.arange(rank, rows, world)
        sel = rrng.permutation(target)[:1500].astype(np.int64)
        src = G.random_table(np.random.default_rng(5000 + ci * 16 + rank), out_dt, sel.shape[0], cols)
        G.scatter(G.np_to_torch(src, out_dt).cuda(), torch.from_numpy(sel).cuda(), t)
        comm.barrier()
        for r in range(world):  # replay every rank's scatter on the host model
            tr = np.arange(r, rows, world)
            rr = np.random.default_rng(77 * ci + r)
            rr.integers(0, rows, size=3000 + r)
            G.random_table(rr, out_dt, 3000 + r, cols)
            s_r = rr.permutation(tr)[:1500].astype(np.int64)
            src_r = G.random_table(np.random.default_rng(5000 + ci * 16 + r), out_dt, s_r.shape[0], cols)
            O.scatter(src_r, out_dt, s_r, host, tab_dt, cols=cols)
        got = G.torch_to_np(local.reshape(nloc, stride), tab_dt)
        assert got.tobytes() == host[first_row:first_row + nloc].tobytes(), f"scatter mismatch case {ci} rank {rank}"
        # and through a remote gather of everything
        all_idx = np.arange(rows, dtype=np.int64)
        full = torch.empty(rows, cols, dtype=G.TORCH_OF[tab_dt], device="cuda")
        G.gather(t, torch.from_numpy(all_idx).cuda(), full)
        torch.cuda.synchronize()
        assert G.torch_to_np(full, tab_dt).tobytes() == np.ascontiguousarray(host[:, :cols]).tobytes(), f"full gather case {ci}"
        comm.barrier()
        wmb.destroy_wholememory_tensor(t)


def scenario_gradient(rank, world, comm):
    import torch
    import wholegraph_b200.torch as wgth
    from oracle import oracle as O
    cfgs = [("adam", {}, "chunked", 392), ("adam", {"adam_w": 1.0, "weight_decay": 0.01}, "distributed", 127),
            ("sgd", {"weight_decay": 0.02}, "continuous", 4), ("adagrad", {"epsilon": 1e-6}, "distributed", 129),
            ("rmsprop", {"alpha": 0.9}, "chunked", 3)]
    rows = 5000
    for ci, (kind, params, mt, dim) in enumerate(cfgs):
        rng = np.random.default_rng(300 + ci)
        w = rng.standard_normal((rows, dim)).astype(np.float32)
        emb = wgth.create_embedding(comm, mt, "cuda", torch.float32, [rows, dim])
        opt = wgth.create_wholememory_optimizer(emb, kind, params, global_comm=comm)
        local, first_row = emb.get_embedding_tensor().get_local_tensor()
        local.copy_(torch.from_numpy(w[first_row:first_row + local.shape[0]]))
        comm.barrier()
        m = np.zeros_like(w)
        v = np.zeros_like(w)
        b12 = np.ones((rows, 2), np.float32)
        for step in range(3):
            all_idx, all_g = [], []
            for r in range(world):
                rr = np.random.default_rng(9000 + ci * 100 + step * 10 + r)
                n = (6000 if (ci == 0 and step == 2) else 700) + 13 * r  # the big step outgrows the first gradient stage
                idx_r = (rr.zipf(1.3, size=n) % rows).astype(np.int64)  # heavy duplication
                g_r = rr.standard_normal((n, dim)).astype(np.float32)
                all_idx.append(idx_r)
                all_g.append(g_r)
            my_idx = all_idx[rank].astype(np.int32 if step == 1 else np.int64)
            emb.add_gradients(torch.from_numpy(my_idx).cuda(), torch.from_numpy(all_g[rank]).cuda())
            emb.need_apply = True
            opt.step(0.01)
            urows, ug = O.dedup_gradients(np.concatenate(all_idx), np.concatenate(all_g))
            kw = dict(weight_decay=params.get("weight_decay", 0.0), epsilon=params.get("epsilon", 1e-8))
            if kind == "adam":
                O.optimizer_step("adam", w, urows, ug, 0.01, state=(m, v), b12=b12, adam_w=params.get("adam_w", 0) > 0.5, **kw)
            elif kind == "sgd":
                O.optimizer_step("sgd", w, urows, ug, 0.01, weight_decay=kw["weight_decay"])
            elif kind == "adagrad":
                O.optimizer_step("adagrad", w, urows, ug, 0.01, state=m, **kw)
            else:
                O.optimizer_step("rmsprop", w, urows, ug, 0.01, state=m, alpha=params.get("alpha", 0.99), **kw)
        torch.cuda.synchronize()
        got = local.cpu().numpy()
        exp = w[first_row:first_row + local.shape[0]]
        assert np.allclose(got, exp, rtol=1e-5, atol=1e-5), f"{kind} mismatch: max abs {np.abs(got - exp).max()}"
        names = emb.get_optimizer_state_names()
        assert names == {"adam": ["m", "v", "beta12t"], "sgd": [], "adagrad": ["state_sum"], "rmsprop": ["v"]}[kind]
        if kind == "adam":
            ml, mfirst = emb.get_optimizer_state("m").get_local_tensor()
            assert np.allclose(ml.cpu().numpy(), m[mfirst:mfirst + ml.shape[0]], rtol=1e-5, atol=1e-6)
        comm.barrier()
        wgth.destroy_wholememory_optimizer(opt)
        wgth.destroy_embedding(emb)


def scenario_sampling(rank, world, comm):
    import torch
    import wholegraph_b200.binding as wmb
    import wholegraph_b200.torch as wgth
    from oracle import oracle as O
    rng = np.random.default_rng(42)
    nodes = 4000
    deg = np.minimum(rng.zipf(1.5, size=nodes), 400) + rng.integers(0, 12, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    for col_dtype, wm_col, th_col in [(np.int64, wmb.DtInt64, torch.int64), (np.int32, wmb.DtInt, torch.int32)]:
        col = rng.integers(0, nodes, size=int(row_ptr[-1])).astype(col_dtype)
        for mt in (wmb.MtChunked, wmb.MtContinuous, wmb.MtDistributed):
            rp = wmb.create_wholememory_array(wmb.DtInt64, nodes + 1, comm.wmb_comm, mt, wmb.MlDevice)
            cp = wmb.create_wholememory_array(wm_col, col.size, comm.wmb_comm, mt, wmb.MlDevice)
            for t, host, dt in ((rp, row_ptr, wmb.DtInt64), (cp, col, wm_col)):
                loc, first = t.get_wholememory_handle().get_local_flatten_tensor(dt, wmb.MlDevice, torch.cuda.current_device())
                loc.copy_(torch.from_numpy(host[first:first + loc.numel()]))
            comm.barrier()
            crng = np.random.default_rng(7 + rank)
            for k, cdt in [(10, np.int64), (25, np.int32), (32, np.int64), (40, np.int64), (100, np.int32), (-1, np.int64)]:
                centers = crng.integers(0, nodes, size=257).astype(cdt)
                seed = 1234 + k
                res = wgth.unweighted_sample_without_replacement(rp, cp, torch.from_numpy(centers).cuda(), k, random_seed=seed,
                                                                 need_center_local_output=True, need_edge_output=True)
                eo, ed, el, eg = O.unweighted_sample(row_ptr, col.astype(np.int64), centers.astype(np.int64), k, seed)
                assert res[0].cpu().numpy().tolist() == eo.tolist(), f"offsets k={k}"
                assert res[1].dtype == th_col and res[1].cpu().numpy().astype(np.int64).tolist() == ed.tolist(), f"dst k={k}"
                assert res[2].cpu().numpy().tolist() == el.tolist() and res[3].cpu().numpy().tolist() == eg.tolist(), f"ids k={k}"
            if mt == wmb.MtChunked and col_dtype == np.int64:
                # two-hop loader step [25, 10] through GraphStructure (sample -> append_unique -> sample), config C5's shape
                gs = wgth.GraphStructure()
                gs.set_csr_graph(wgth.WholeMemoryTensor(rp), wgth.WholeMemoryTensor(cp))
                seeds = torch.from_numpy(np.random.default_rng(99 + rank).permutation(nodes)[:64].astype(np.int64)).cuda()
                tg, ei, rps, cis = gs.multilayer_sample_without_replacement(seeds, [25, 10], random_seed=4242)
                assert len(tg) == 3 and torch.equal(tg[2], seeds)
                for hop, (k, seed) in enumerate([(10, 4242 + 0), (25, 4242 + 1)]):
                    centers_h = tg[hop + 1].cpu().numpy()
                    eo, ed, el, _ = O.unweighted_sample(row_ptr, col.astype(np.int64), centers_h, k, seed)
                    assert rps[hop].cpu().numpy().tolist() == eo.tolist()
                    uniq = tg[hop].cpu().numpy()
                    assert np.array_equal(uniq[:centers_h.size], centers_h)            # targets stay in front
                    assert np.array_equal(uniq[cis[hop].cpu().numpy()], ed)            # sampled dst ids, via the unique map
                    assert np.array_equal(ei[hop][1].cpu().numpy(), el)                # source local ids
                    assert len(set(uniq.tolist())) == uniq.size
            comm.barrier()
            wmb.destroy_wholememory_tensor(rp)
            wmb.destroy_wholememory_tensor(cp)


def scenario_file_io(rank, world, comm):
    """Round trip through the reference's on-disk format (raw row-major part files "<prefix>_part_i_of_n"):
    store per-rank parts, reload them into differently typed / partitioned tables, and reload files whose part
    count differs from the world size."""
    import tempfile
    import torch
    import wholegraph_b200.torch as wgth
    from oracle import oracle as O
    rows, cols, stride = 3001, 13, 16
    rng = np.random.default_rng(77)
    host = rng.standard_normal((rows, stride)).astype(np.float32)
    base = os.path.join(tempfile.gettempdir(), "wgb200_io_%s" % os.environ["MASTER_PORT"])
    # 1. files written by "someone else": 3 part files of the [rows, cols] matrix, loaded by `world` ranks
    if rank == 0:
        cuts = [0, 1000, 1001, rows]
        for i in range(3):
            host[cuts[i]:cuts[i + 1], :cols].copy().tofile("%s_src_part_%d_of_3" % (base, i))
    comm.barrier()
    filelist = ["%s_src_part_%d_of_3" % (base, i) for i in range(3)]
    for mt, loc in (("chunked", "cuda"), ("distributed", "cuda"), ("continuous", "cpu")):
        t = wgth.create_wholememory_tensor(comm, mt, loc, [rows, cols], torch.float32, [stride, 1])
        t.from_filelist(filelist)
        local, first = t.get_local_tensor(host_view=(loc == "cpu"))
        assert np.array_equal(local.cpu().numpy(), host[first:first + local.shape[0], :cols]), (mt, loc)
        idx = torch.from_numpy(rng.integers(0, rows, size=500).astype(np.int64)).cuda()
        assert np.array_equal(t.gather(idx).cpu().numpy(), host[idx.cpu().numpy(), :cols])
        # 2. store my part, then reload everything into a table of another type
        t.to_file_prefix(base + "_dump_" + mt)
        comm.barrier()
        t2 = wgth.create_wholememory_tensor_from_filelist(comm, "continuous", "cuda",
                                                          ["%s_dump_%s_part_%d_of_%d" % (base, mt, r, world) for r in range(world)],
                                                          torch.float32, last_dim_size=cols)
        assert t2.shape == (rows, cols)
        l2, f2 = t2.get_local_tensor()
        assert np.array_equal(l2.cpu().numpy(), host[f2:f2 + l2.shape[0], :cols])
        comm.barrier()
        wgth.destroy_wholememory_tensor(t2)
        wgth.destroy_wholememory_tensor(t)
    # 3. embedding + optimizer state checkpoint round trip
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, cols])
    opt = wgth.create_wholememory_optimizer(emb, "adam", {}, global_comm=comm)
    lw, fw = emb.get_embedding_tensor().get_local_tensor()
    lw.copy_(torch.from_numpy(host[fw:fw + lw.shape[0], :cols]))
    comm.barrier()
    gi = torch.from_numpy(np.random.default_rng(5 + rank).integers(0, rows, size=400).astype(np.int64)).cuda()
    if os.environ.get("WG_TEST_SHARED_GPU") != "1":  # the gradient exchange needs NCCL = one GPU per rank
        emb.add_gradients(gi, torch.ones(400, cols, device="cuda"))
        emb.need_apply = True
        opt.step(0.1)
    emb.save(base + "_ckpt")
    comm.barrier()
    emb2 = wgth.create_embedding(comm, "distributed", "cuda", torch.float32, [rows, cols])
    opt2 = wgth.create_wholememory_optimizer(emb2, "adam", {}, global_comm=comm)
    emb2.load(base + "_ckpt")
    for get in (lambda e: e.get_embedding_tensor(), lambda e: e.get_optimizer_state("m"), lambda e: e.get_optimizer_state("v"),
                lambda e: e.get_optimizer_state("beta12t")):
        a, _ = get(emb).get_local_tensor()
        b, _ = get(emb2).get_local_tensor()
        assert torch.equal(a, b)
    comm.barrier()
    for o, e in ((opt, emb), (opt2, emb2)):
        wgth.destroy_wholememory_optimizer(o)
        wgth.destroy_embedding(e)
    if rank == 0:
        import glob
        for f in glob.glob(base + "_*"):
            os.remove(f)


def scenario_file_io_grid(rank, world, comm):
    """The reference's load / store test grid (python/pylibwholegraph/.../tests/pylibwholegraph/test_wholememory_io.py:177-186,
    :344-351; round_robin_size 0 only): part-file counts {3, 5}, embedding dim {16, 31, 33} inside row strides {32, 64},
    storage offsets {0, 3} (a column window of a wider table), default / random row partitions, every memory type.  After a
    load the window holds the file rows and the columns outside it keep their sentinel; a store writes exactly the window."""
    import tempfile
    import torch
    import wholegraph_b200.torch as wgth
    rows = 1537
    base = os.path.join(tempfile.gettempdir(), "wgb200_iogrid_%s" % os.environ["MASTER_PORT"])
    cases = [(3, 16, 32, 0, "default", "chunked", "cuda"), (5, 31, 32, 0, "random", "continuous", "cuda"),
             (3, 33, 64, 3, "random", "distributed", "cuda"), (5, 16, 64, 3, "default", "chunked", "cpu"),
             (3, 31, 64, 3, "random", "continuous", "cpu"), (5, 33, 64, 0, "default", "chunked", "cuda")]
    for ci, (parts, dim, stride, off, method, mt, loc) in enumerate(cases):
        rng = np.random.default_rng(4000 + ci)  # same stream on every rank
        host = rng.standard_normal((rows, dim)).astype(np.float32)
        cuts = [0] + sorted(rng.choice(np.arange(1, rows), size=parts - 1, replace=False).tolist()) + [rows]
        partition = _random_partition(rng, rows, world) if (method == "random" and world > 1) else None
        files = ["%s_c%d_part_%d_of_%d" % (base, ci, i, parts) for i in range(parts)]
        if rank == 0:
            for i, f in enumerate(files):
                host[cuts[i]:cuts[i + 1]].copy().tofile(f)
        comm.barrier()
        root = wgth.create_wholememory_tensor(comm, mt, loc, [rows, stride], torch.float32, [stride, 1], partition)
        whole_local, first = root.get_local_tensor(host_view=(loc == "cpu"))
        whole_local.fill_(-7.0)
        comm.barrier()
        window = root.get_sub_tensor([0, off], [rows, off + dim])
        assert tuple(window.shape) == (rows, dim) and window.stride()[0] == stride and window.storage_offset() == off
        window.from_filelist(files)
        comm.barrier()
        mine = whole_local.cpu().numpy()
        nloc = mine.shape[0]
        if partition is not None:
            assert nloc == partition[rank] and first == sum(partition[:rank])
        assert np.array_equal(mine[:, off:off + dim], host[first:first + nloc]), ("load", ci)
        outside = np.delete(mine, np.s_[off:off + dim], axis=1)
        assert np.all(outside == -7.0), ("load wrote outside the column window", ci)
        # store: every rank writes the window of its rows; the concatenation of the parts is the file content again
        window.to_file_prefix("%s_c%d_out" % (base, ci))
        comm.barrier()
        if rank == 0:
            back = np.concatenate([np.fromfile("%s_c%d_out_part_%d_of_%d" % (base, ci, r, world), dtype=np.float32) for r in range(world)])
            assert back.size == rows * dim and np.array_equal(back.reshape(rows, dim), host), ("store", ci)
        comm.barrier()
        wgth.destroy_wholememory_tensor(root)
    if rank == 0:
        import glob
        for f in glob.glob(base + "_*"):
            os.remove(f)


def scenario_sampling_grid(rank, world, comm):
    """The reference's Python sampling test grids (tests/wholegraph_torch/ops/test_wholegraph_unweighted_sample_without_replacement.py
    :360-369, ..._weighted_...:362-372): graph of 103 / 113 nodes and 1043 edges, 13 centers, fan-out 11 and -1, int32 / int64
    centers and column ids, CSR in DEVICE and in HOST memory, CONTINUOUS / CHUNKED / DISTRIBUTED, every combination of the two
    optional outputs -- the number and order of returned tensors follow the flags."""
    import torch
    import wholegraph_b200.binding as wmb
    import wholegraph_b200.torch as wgth
    from oracle import oracle as O
    dev = torch.cuda.current_device()

    def make_graph(rng, nodes, edges):
        deg = rng.multinomial(edges, np.ones(nodes) / nodes)
        row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        return row_ptr

    def fill(t, host, dt, ml):
        loc_t, first = t.get_wholememory_handle().get_local_flatten_tensor(dt, ml, dev if ml == wmb.MlDevice else -1)
        loc_t.copy_(torch.from_numpy(host[first:first + loc_t.numel()]))

    # ---- unweighted
    rng = np.random.default_rng(360)
    nodes, edges = 103, 1043
    row_ptr = make_graph(rng, nodes, edges)
    for col_np, col_wm in ((np.int32, wmb.DtInt), (np.int64, wmb.DtInt64)):
        col = rng.integers(0, nodes, size=edges).astype(col_np)
        for ml in (wmb.MlDevice, wmb.MlHost):
            for mt in (wmb.MtContinuous, wmb.MtChunked, wmb.MtDistributed):
                if ml == wmb.MlHost and mt == wmb.MtDistributed and world > 1:
                    continue  # host DISTRIBUTED shards are private to their rank (no mapping): exchange path, covered by the device case
                rp = wmb.create_wholememory_array(wmb.DtInt64, nodes + 1, comm.wmb_comm, mt, ml)
                cp = wmb.create_wholememory_array(col_wm, edges, comm.wmb_comm, mt, ml)
                fill(rp, row_ptr, wmb.DtInt64, ml)
                fill(cp, col, col_wm, ml)
                comm.barrier()
                for k in (11, -1):
                    for cen_np in (np.int32, np.int64):
                        centers = np.random.default_rng(13 + rank).integers(0, nodes, size=13).astype(cen_np)
                        eo, ed, el, eg = O.unweighted_sample(row_ptr, col.astype(np.int64), centers.astype(np.int64), k, 1000 + k)
                        for need_lid in (True, False):
                            for need_gid in (True, False):
                                res = wgth.unweighted_sample_without_replacement(rp, cp, torch.from_numpy(centers).cuda(), k, random_seed=1000 + k,
                                                                                 need_center_local_output=need_lid, need_edge_output=need_gid)
                                assert len(res) == 2 + int(need_lid) + int(need_gid), (len(res), need_lid, need_gid)
                                assert res[0].dtype == torch.int32 and res[0].cpu().numpy().tolist() == eo.tolist()
                                assert res[1].cpu().numpy().astype(np.int64).tolist() == ed.tolist(), (k, ml, mt)
                                pos = 2
                                if need_lid:
                                    assert res[pos].cpu().numpy().tolist() == el.tolist()
                                    pos += 1
                                if need_gid:
                                    assert res[pos].dtype == torch.int64 and res[pos].cpu().numpy().tolist() == eg.tolist()
                comm.barrier()
                wmb.destroy_wholememory_tensor(rp)
                wmb.destroy_wholememory_tensor(cp)

    # ---- weighted
    rng = np.random.default_rng(362)
    nodes, edges = 113, 1043
    row_ptr = make_graph(rng, nodes, edges)
    for col_np, col_wm in ((np.int32, wmb.DtInt), (np.int64, wmb.DtInt64)):
        col = rng.integers(0, nodes, size=edges).astype(col_np)
        for w_np, w_wm in ((np.float32, wmb.DtFloat), (np.float64, wmb.DtDouble)):
            weights = rng.uniform(0.1, 3.0, size=edges).astype(w_np)
            for ml in (wmb.MlDevice, wmb.MlHost):
                for mt in (wmb.MtContinuous, wmb.MtChunked):
                    rp = wmb.create_wholememory_array(wmb.DtInt64, nodes + 1, comm.wmb_comm, mt, ml)
                    cp = wmb.create_wholememory_array(col_wm, edges, comm.wmb_comm, mt, ml)
                    wp = wmb.create_wholememory_array(w_wm, edges, comm.wmb_comm, mt, ml)
                    fill(rp, row_ptr, wmb.DtInt64, ml)
                    fill(cp, col, col_wm, ml)
                    fill(wp, weights, w_wm, ml)
                    comm.barrier()
                    for cen_np in (np.int32, np.int64):
                        centers = np.random.default_rng(17 + rank).integers(0, nodes, size=13).astype(cen_np)
                        eo, ed, el, eg, margin = O.weighted_sample(row_ptr, col.astype(np.int64), weights, centers.astype(np.int64), 11, 77)
                        for need_lid in (True, False):
                            for need_gid in (True, False):
                                res = wgth.weighted_sample_without_replacement(rp, cp, wp, torch.from_numpy(centers).cuda(), 11, random_seed=77,
                                                                               need_center_local_output=need_lid, need_edge_output=need_gid)
                                assert len(res) == 2 + int(need_lid) + int(need_gid)
                                off = res[0].cpu().numpy()
                                assert off.tolist() == eo.tolist()
                                dst = res[1].cpu().numpy().astype(np.int64)
                                for c in range(centers.size):  # per-center sets, the reference's own comparison (segment_sort_output)
                                    a, b = off[c], off[c + 1]
                                    if sorted(dst[a:b].tolist()) != sorted(ed[a:b].tolist()):
                                        assert margin[c] < 1e-5, (c, margin[c])
                                if need_lid:
                                    assert res[2].cpu().numpy().tolist() == el.tolist()
                                if need_gid:
                                    gid = res[2 + int(need_lid)].cpu().numpy()
                                    assert np.array_equal(col[gid].astype(np.int64), dst)  # edge ids name the sampled neighbours
                    comm.barrier()
                    for t in (rp, cp, wp):
                        wmb.destroy_wholememory_tensor(t)


def scenario_gather_scatter_functors(rank, world, comm):
    """The reference's torch-level op test (tests/wholegraph_torch/ops/test_wholegraph_gather_scatter.py:40-160) with its own
    sizes: a [1024*256*world + 3, 256] fp32 table under a random row partition, every rank scatters the rows
    rank, rank + world, ... through wholememory_scatter_functor, the local view is compared with the closed form
    value(row, col) = row + col, then 100,001 random int32 indices are gathered with wholememory_gather_forward_functor.
    Every memory type x location the communicator supports."""
    import torch
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch import wholememory_ops as wm_ops
    from wholegraph_b200.torch.dlpack_utils import torch_import_from_dlpack
    count, dim, n = 1024 * 256 * world + 3, 256, 100001
    partition = _random_partition(np.random.default_rng(1), count, world) if world > 1 else None
    dev = torch.cuda.current_device()

    def closed_form(ids):
        return (ids.to(torch.int64).reshape(-1, 1) + torch.arange(dim, dtype=torch.int64).reshape(1, dim)).to(torch.float32)

    for mt in (wmb.MtContinuous, wmb.MtChunked, wmb.MtDistributed):
        for ml in (wmb.MlHost, wmb.MlDevice):
            if not comm.wmb_comm.support_type_location(mt, ml):
                continue
            if mt == wmb.MtDistributed and ml == wmb.MlHost and world > 1 and os.environ.get("WG_TEST_SHARED_GPU") == "1":
                continue  # private host shards travel over NCCL, which needs one GPU per rank (covered by test_one_rank_per_gpu)
            table = wmb.create_wholememory_matrix(wmb.DtFloat, count, dim, -1, comm.wmb_comm, mt, ml, partition)
            mine = torch.arange(rank, count, world, dtype=torch.int64)
            wm_ops.wholememory_scatter_functor(closed_form(mine).cuda(), mine.cuda(), table)
            torch.cuda.synchronize()
            comm.barrier()
            local, start = table.get_local_tensor(torch_import_from_dlpack, wmb.MlDevice, dev)  # the reference's call form
            assert start == table.get_local_entry_start() and tuple(local.shape) == (table.get_local_entry_count(), dim)
            if partition is not None:
                assert table.get_local_entry_count() == partition[rank]
            assert torch.equal(local.cpu(), closed_form(torch.arange(start, start + local.shape[0])))
            idx = torch.randint(0, count, (n,), dtype=torch.int32, generator=torch.Generator().manual_seed(5 + rank))
            rows = wm_ops.wholememory_gather_forward_functor(table, idx.cuda())
            assert rows.dtype == torch.float32 and torch.equal(rows.cpu(), closed_form(idx))
            half = wm_ops.wholememory_gather_forward_functor(table, idx.cuda()[:1000], torch_output_dtype=torch.float16)
            assert half.dtype == torch.float16 and torch.equal(half.cpu(), closed_form(idx[:1000]).to(torch.float16))
            comm.barrier()
            del local
            wmb.destroy_wholememory_tensor(table)


def scenario_weighted_sampling(rank, world, comm):
    """Weighted (A-Res) sampler vs the oracle: same sample sets for the same seed.  Keys are float log1pf/exp2f values, so a
    center whose k-th and (k+1)-th keys are within a few ulp may legitimately resolve differently between libm and CUDA:
    those (flagged by the oracle's margin) are only required to be valid samples."""
    import torch
    import wholegraph_b200.binding as wmb
    import wholegraph_b200.torch as wgth
    from oracle import oracle as O
    rng = np.random.default_rng(4242)
    nodes = 3000
    deg = np.minimum(rng.zipf(1.4, size=nodes), 700) + rng.integers(0, 20, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    edges = int(row_ptr[-1])
    col = rng.integers(0, nodes, size=edges).astype(np.int64)
    for w_np, w_wm in ((np.float32, wmb.DtFloat), (np.float64, wmb.DtDouble)):
        weights = rng.uniform(0.05, 4.0, size=edges).astype(w_np)
        for mt in (wmb.MtChunked, wmb.MtContinuous):
            rp = wmb.create_wholememory_array(wmb.DtInt64, nodes + 1, comm.wmb_comm, mt, wmb.MlDevice)
            cp = wmb.create_wholememory_array(wmb.DtInt64, edges, comm.wmb_comm, mt, wmb.MlDevice)
            wp = wmb.create_wholememory_array(w_wm, edges, comm.wmb_comm, mt, wmb.MlDevice)
            for t, host, dt in ((rp, row_ptr, wmb.DtInt64), (cp, col, wmb.DtInt64), (wp, weights, w_wm)):
                loc, first = t.get_wholememory_handle().get_local_flatten_tensor(dt, wmb.MlDevice, torch.cuda.current_device())
                loc.copy_(torch.from_numpy(host[first:first + loc.numel()]))
            comm.barrier()
            crng = np.random.default_rng(11 + rank)
            for k in (5, 10, 25, 32, 40, 300, -1):
                centers = crng.integers(0, nodes, size=301).astype(np.int64)
                seed = 99 + k
                res = wgth.weighted_sample_without_replacement(rp, cp, wp, torch.from_numpy(centers).cuda(), k, random_seed=seed,
                                                               need_center_local_output=True, need_edge_output=True)
                eo, ed, el, eg, margin = O.weighted_sample(row_ptr, col, weights, centers, k, seed)
                off = res[0].cpu().numpy()
                assert off.tolist() == eo.tolist(), f"offsets k={k}"
                dst, lid, gid = res[1].cpu().numpy(), res[2].cpu().numpy(), res[3].cpu().numpy()
                assert lid.tolist() == el.tolist()
                loose = 0
                for c in range(centers.size):
                    s, e = off[c], off[c + 1]
                    lo, hi = row_ptr[centers[c]], row_ptr[centers[c] + 1]
                    assert np.all((gid[s:e] >= lo) & (gid[s:e] < hi)) and len(set(gid[s:e].tolist())) == e - s
                    assert np.array_equal(dst[s:e], col[gid[s:e]])
                    if set(gid[s:e].tolist()) != set(eg[s:e].tolist()):
                        assert margin[c] < 1e-5, f"center {c} (k={k}): different sample set with key margin {margin[c]}"
                        assert len(set(gid[s:e].tolist()) ^ set(eg[s:e].tolist())) <= 2
                        loose += 1
                assert loose <= 3, f"{loose} centers resolved near-ties differently (k={k})"
            comm.barrier()
            for t in (rp, cp, wp):
                wmb.destroy_wholememory_tensor(t)


def scenario_failing_together(rank, world, comm):
    """A rank-local failure inside a collective call reaches every rank as an error instead of leaving the others in a
    rendezvous: (1) one rank names a part file that does not exist, (2) the ranks pass different row partitions with the
    same total, (3) the communicator still works afterwards."""
    import tempfile
    import torch
    import wholegraph_b200.binding as wmb
    rows, dim = 4096, 16
    t = wmb.create_wholememory_matrix(wmb.DtFloat, rows, dim, -1, comm.wmb_comm, wmb.MtChunked, wmb.MlDevice)
    tmp = os.path.join(tempfile.gettempdir(), "wg_failing_together_%d" % os.getppid())
    if rank == 0:
        np.arange(rows * dim, dtype=np.float32).tofile(tmp)
    comm.barrier()
    names = [tmp if rank != world - 1 else tmp + ".does-not-exist"]
    try:
        t.from_filelist(names)
        raised = None
    except Exception as e:  # the rank with the missing file reports INVALID_INPUT, the others learn that it failed
        raised = e
    assert raised is not None, "rank %d: load with a missing file on rank %d did not fail here" % (rank, world - 1)
    t.from_filelist([tmp])  # same call, good arguments: works, so nobody is out of step
    local, first = t.get_local_tensor(wmb.MlDevice, torch.cuda.current_device())
    assert torch.equal(local.cpu().reshape(-1), torch.arange(first * dim, (first + local.shape[0]) * dim, dtype=torch.float32))
    comm.barrier()
    wmb.destroy_wholememory_tensor(t)
    if world > 1:
        part = [rows // world] * world
        part[-1] += rows - sum(part)
        if rank == 1:  # same total, different split
            part[0] -= 1
            part[-1] += 1
        try:
            bad = wmb.create_wholememory_matrix(wmb.DtFloat, rows, dim, -1, comm.wmb_comm, wmb.MtChunked, wmb.MlDevice, part)
            wmb.destroy_wholememory_tensor(bad)
            raise AssertionError("ranks with different partitions built a table")
        except RuntimeError as e:
            assert "ogic" in str(e), e  # WHOLEMEMORY_LOGIC_ERROR on every rank
    comm.barrier()
    t2 = wmb.create_wholememory_matrix(wmb.DtFloat, rows, dim, -1, comm.wmb_comm, wmb.MtContinuous, wmb.MlDevice)
    wmb.destroy_wholememory_tensor(t2)
    if rank == 0:
        os.unlink(tmp)


SCENARIOS = {"failing_together": scenario_failing_together, "gather_scatter_functors": scenario_gather_scatter_functors, "sampling_grid": scenario_sampling_grid, "file_io_grid": scenario_file_io_grid, "weighted_sampling": scenario_weighted_sampling, "file_io": scenario_file_io, "gather_scatter": scenario_gather_scatter, "gradient": scenario_gradient, "sampling": scenario_sampling}


def worker(rank, world, port, ngpus, scenario, env, results):
    try:
        comm = _setup(rank, world, port, ngpus, env)
        SCENARIOS[scenario](rank, world, comm)
        results[rank] = "ok"
    except Exception:
        results[rank] = "FAIL: " + traceback.format_exc()
    finally:
        try:
            import wholegraph_b200.torch as wgth
            wgth.finalize()
        except Exception:
            pass
