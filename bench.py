#!/usr/bin/env python
"""bench.py -- device-timed WholeMemory embedding gather (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                      # N=1: configs[1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               # the reference's own kernels (oracle/_ref)

A "step" is ONE wholememory_gather call over one batch of 1,048,576 uniform-random int64 indices.

N=1 workload = BASELINE.json configs[1]: CONTINUOUS/DEVICE table 100M x 256 fp32 (102.4 GB) on one B200.
N>1 (weak scaling): CHUNKED/DEVICE table of N x 100M rows x 256 fp32 row-sharded over the N GPUs of the box
and mapped into every GPU (VMM over NVSwitch); every rank gathers 1M rows drawn uniformly from the WHOLE table.

value       = n * row_out_bytes * N / t   (GB/s of gathered output -- the reference bench's "Bandwidth",
              cpp/bench/wholememory_ops/gather_scatter_bench.cu:363-366), t = max over ranks of the CUDA-event
              time of the K timed calls / K, inputs resident in HBM.
e2e         = same metric with the step's indices starting in pinned HOST memory (H2D inside the timed region)
              and the gathered rows copied back to pinned host memory (D2H inside the timed region).
roofline    = algorithmic bytes n*(row_in + row_out + idx) / kernel time against the measured HBM copy
              bandwidth (N=1) -- for N>1 also the NVLink-ingress bound is reported.
cpu_baseline= the oracle's multithreaded host gather (oracle/wm_oracle.c: oracle_gather_mt) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 100_000_000
DIM = 256
BATCH = 1 << 20
METRIC = "embedding gather GB/s (device-timed)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    p.add_argument("--dim", type=int, default=DIM)
    p.add_argument("--batch", type=int, default=BATCH)
    p.add_argument("--dtype", default="fp32", choices=["fp32", "fp16"])
    p.add_argument("--memory-type", default=None, choices=[None, "continuous", "chunked", "distributed"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--index-pattern", default="random", choices=["random", "sequential", "remote", "local"],
                   help="diagnostic; the metric is 'random' (uniform over the whole table)")
    p.add_argument("--per-step-events", action="store_true", help="diagnostic: also time every step with its own event pair")
    return p.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = self.samples[0][1]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(mx)) if mx.replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(dim, esize, np_dtype, budget_s=12.0):
    """Oracle multithreaded host gather on a bounded sample of the same workload (rows x dim, 1M-index batches)."""
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    rows = 4_000_000 if esize == 4 else 8_000_000  # ~4 GB table, far beyond any LLC
    pinned = False
    try:  # north_star asks for a pinned-host table; page-locking does not change what the host cores see
        import torch
        if torch.cuda.is_available():
            th = torch.empty((rows, dim), dtype=torch.float32 if esize == 4 else torch.float16, pin_memory=True)
            table, pinned = th.numpy(), True
    except Exception:
        pinned = False
    if not pinned:
        try:
            table = np.empty((rows, dim), dtype=np_dtype)
        except MemoryError:
            rows //= 4
            table = np.empty((rows, dim), dtype=np_dtype)
    table[:] = (np.arange(rows, dtype=np.int64) & 0xFFFF).astype(np_dtype)[:, None]
    rng = np.random.default_rng(0x5EED)
    n = BATCH
    out = np.empty((n, dim), dtype=np_dtype)
    idx = rng.integers(0, rows, size=n).astype(np.int64)
    O.gather_mt(table, idx, out, cores)  # warm-up + page-in
    assert np.array_equal(out[:, 0], table[idx, 0])
    best, passes, t_total = 0.0, 0, 0.0
    while t_total < budget_s and passes < 200:
        idx = rng.integers(0, rows, size=n).astype(np.int64)
        t0 = time.perf_counter()
        O.gather_mt(table, idx, out, cores)
        dt = time.perf_counter() - t0
        t_total += dt
        passes += 1
        best = max(best, n * dim * esize / dt / 1e9)
    numpy_take = None
    try:  # single-thread numpy.take as the "numpy-index reference"
        t0 = time.perf_counter()
        np.take(table, idx, axis=0, out=out)
        numpy_take = round(n * dim * esize / (time.perf_counter() - t0) / 1e9, 3)
    except Exception:
        pass
    return {"value": round(best, 3), "unit": "GB/s", "cores": cores, "kind": "port", "numpy_take_1_thread_gbs": numpy_take,
            "sample": "%d passes of %d random rows x %d B from a %d-row %s host table (oracle_gather_mt, %d threads, best pass)"
                      % (passes, n, dim * esize, rows, "pinned (cudaHostAlloc)" if pinned else "pageable", cores)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # the reference library is loaded THROUGH THE SAME BINDING (identical C ABI): oracle/_ref build
        ref = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")
        if os.path.exists(ref):
            os.environ["WHOLEGRAPH_B200_LIB"] = ref
        else:
            return reference_cpu_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_wholegraph_env_fns, wrap_torch_tensor

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
    uid = wmb.create_unique_id() if rank == 0 else wmb.PyWholeMemoryUniqueID()
    if world > 1:
        t = uid.as_tensor().cuda()
        dist.broadcast(t, 0)
        uid.as_tensor().copy_(t.cpu())
    comm = wmb.create_communicator(uid, rank, world)

    th_dtype, wm_dtype, esize = (torch.float32, wmb.DtFloat, 4) if args.dtype == "fp32" else (torch.float16, wmb.DtHalf, 2)
    dim, n = args.dim, args.batch
    rows_total = args.rows_per_gpu * world
    mem_type = args.memory_type or ("continuous" if world == 1 else "chunked")
    mt = {"continuous": wmb.MtContinuous, "chunked": wmb.MtChunked, "distributed": wmb.MtDistributed}[mem_type]
    table = wmb.create_wholememory_matrix(wm_dtype, rows_total, dim, -1, comm, mt, wmb.MlDevice)
    # fill my shard with the reference tests' closed-form pattern (row id & mask) so any gathered row is checkable
    local, first_row = table.get_local_tensor(wmb.MlDevice, local_rank)
    mask = (1 << 24) - 1 if esize == 4 else (1 << 11) - 1
    chunk = 4_000_000
    for s in range(0, local.shape[0], chunk):
        e = min(local.shape[0], s + chunk)
        ids = torch.arange(first_row + s, first_row + e, device="cuda", dtype=torch.int64)
        local[s:e] = (ids & mask).to(th_dtype).unsqueeze(1)
    torch.cuda.synchronize()
    comm.barrier()

    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED + rank)
    n_batches = 8  # distinct index batches, cycled: successive steps never re-read the same rows
    idx_dev = [torch.randint(0, rows_total, (n,), device="cuda", dtype=torch.int64, generator=gen) for _ in range(n_batches)]
    if args.index_pattern in ("remote", "local") and world > 1:  # diagnostic: only peer rows / only my rows
        per = args.rows_per_gpu
        idx_dev = []
        for _ in range(n_batches):
            r = torch.randint(0, per * (world - 1) if args.index_pattern == "remote" else per, (n,), device="cuda", dtype=torch.int64, generator=gen)
            if args.index_pattern == "remote":
                r = r + (r >= rank * per).to(torch.int64) * per  # skip my own partition
            else:
                r = r + rank * per
            idx_dev.append(r)
    if args.index_pattern == "sequential":  # diagnostic only: contiguous rows = the kernel's ceiling without DRAM page misses
        idx_dev = [(torch.arange(n, device="cuda", dtype=torch.int64) + (b * n * 7) % max(1, rows_total - n)) for b in range(n_batches)]
    out = torch.empty(n, dim, device="cuda", dtype=th_dtype)
    env = get_wholegraph_env_fns()
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    w_idx = [wrap_torch_tensor(t) for t in idx_dev]
    w_out = wrap_torch_tensor(out)

    def step(i):
        wmb.wholememory_gather_op(table, w_idx[i % n_batches], w_out, env, sptr)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    # correctness of the timed configuration (cheap, outside the timed region)
    torch.cuda.synchronize()
    last = (max(args.warmup, 3) - 1) % n_batches
    assert torch.equal(out[:, 0], (idx_dev[last] & mask).to(th_dtype)) and torch.equal(out[:, dim - 1], out[:, 0]), "gather wrong"

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # the reference bench's own clock: wall time over K calls + one sync
    sync_all()
    ms = ev0.elapsed_time(ev1) / args.steps
    # keep the GPU busy a little longer so the clock sampler sees the loaded state even for tiny K
    t_end = time.time() + 1.0
    k = 0
    while time.time() < t_end:
        step(k)
        k += 1
    torch.cuda.synchronize()
    clocks = sampler.summary() if sampler else None

    if args.per_step_events and rank == 0:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i, (a, b) in enumerate(evs):
            a.record(stream)
            step(i)
            b.record(stream)
        torch.cuda.synchronize()
        per = sorted(a.elapsed_time(b) for a, b in evs)
        print("per-step kernel ms: min %.4f median %.4f max %.4f (loop mean %.4f)" % (per[0], per[len(per) // 2], per[-1], ms), file=sys.stderr)
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())

    # ---- end to end: indices from pinned host, rows back to pinned host, both copies inside the timed region
    e2e = None
    if not args.no_e2e:
        idx_host = [t.cpu().pin_memory() for t in idx_dev[:2]]
        out_host = torch.empty(n, dim, dtype=th_dtype).pin_memory()
        idx_stage = torch.empty(n, device="cuda", dtype=torch.int64)
        w_stage = wrap_torch_tensor(idx_stage)
        e_steps = max(3, min(args.steps, 10))

        def e2e_step(i):
            idx_stage.copy_(idx_host[i % 2], non_blocking=True)
            wmb.wholememory_gather_op(table, w_stage, w_out, env, sptr)
            out_host.copy_(out, non_blocking=True)

        e2e_step(0)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(e_steps):
            e2e_step(i)
        e1.record(stream)
        sync_all()
        e_ms = torch.tensor([e0.elapsed_time(e1) / e_steps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        assert torch.equal(out_host[:, 0], (idx_host[(e_steps - 1) % 2] & mask).to(th_dtype))
        e2e = {"value": round(n * dim * esize * world / (float(e_ms.item()) * 1e-3) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * dim * esize, "ms_per_step": round(float(e_ms.item()), 4),
               "note": "PCIe-bound: the %d MB of gathered rows cross PCIe to pinned host memory every step" % (n * dim * esize >> 20)}
        # Informational only (NOT the e2e figure): the way the library is used by a GNN loader -- indices arrive from pinned
        # host memory, the gathered rows stay in HBM for the model, and 8 bytes (a checksum of the step's rows) return.
        try:
            if world > 1:
                raise RuntimeError("single-GPU leg only (a rank-local failure must not strand the others in a collective)")
            chk_host = torch.empty(1, dtype=torch.float64).pin_memory()

            def resident_step(i):
                idx_stage.copy_(idx_host[i % 2], non_blocking=True)
                wmb.wholememory_gather_op(table, w_stage, w_out, env, sptr)
                chk_host.copy_(out[:, 0].sum(dtype=torch.float64).reshape(1), non_blocking=True)

            resident_step(0)
            sync_all()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for i in range(e_steps):
                resident_step(i)
            r1.record(stream)
            sync_all()
            r_ms = torch.tensor([r0.elapsed_time(r1) / e_steps], device="cuda", dtype=torch.float64)
            exp_chk = float((idx_host[(e_steps - 1) % 2] & mask).to(torch.float64).sum())
            if abs(float(chk_host.item()) - exp_chk) <= 1e-6 * max(1.0, abs(exp_chk)):
                e2e["device_resident_output"] = {"value": round(n * dim * esize * world / (float(r_ms.item()) * 1e-3) / 1e9, 3), "unit": "GB/s",
                                                 "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": 8,
                                                 "ms_per_step": round(float(r_ms.item()), 4),
                                                 "note": "informational: rows stay in HBM, an 8-byte checksum returns"}
        except Exception as ex:  # never let the informational leg break the bench line
            if world == 1:
                print("device-resident e2e leg skipped: %r" % (ex,), file=sys.stderr)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        row = dim * esize
        out_gbs = n * row * world / (ms_max * 1e-3) / 1e9
        alg_bytes = n * (row + row + 8)
        achieved = alg_bytes / (ms_max * 1e-3) / 1e9  # per GPU
        t_hbm = alg_bytes / (hbm_peak * 1e9)
        t_nvl = n * row * (world - 1) / world / 770e9  # measured peer-copy bandwidth per direction (B200_PROFILING.md)
        bound_ms = max(t_hbm, t_nvl) * 1e3
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the committed `ncu --set full` capture
        # (profiles/r1_gather_c2_full_summary.txt: same command, same kernel configuration); not re-measured in this run.
        traffic, traffic_src = None, None
        if args.impl != "reference" and world == 1 and (dim, esize, n) == (256, 4, 1 << 20):
            traffic, traffic_src = 2.100e9, "profiles/r1_gather_c2_full_summary.txt (ncu --set full, 1.081 GB read + 1.020 GB written)"
        roof = {"bound": "hbm" if t_hbm >= t_nvl else "nvlink", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "wm::row_move_vec_kernel<int64,16B,gather>", "kernel_ms": round(ms_max, 4),
                "bound_ms": round(bound_ms, 4), "frac_of_bound_time": round(bound_ms / ms_max, 4),
                "algorithmic_bytes_per_launch": alg_bytes}
        line = {
            "metric": METRIC, "value": round(out_gbs, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if esize == 4 else "f16", "data": "synthetic",
            "config": {"workload": "%s/DEVICE gather, %d x %d %s table (%.1f GB per GPU), %d uniform-random int64 indices per rank per step"
                                   % (mem_type.upper(), rows_total, dim, args.dtype, args.rows_per_gpu * row / 1e9, n),
                       "l2": "inputs larger than L2: table %.0f GB, 8 distinct index batches cycled, 1 GB output per step" % (rows_total * row / 1e9),
                       "rows_per_gpu": args.rows_per_gpu, "embedding_dim": dim, "indices_per_rank": n, "memory_type": mem_type},
            "roofline": roof, "gpu_launches": args.steps, "clocks": clocks,
            "host_timed": {"value": round(n * row * world / (host_ms * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(host_ms, 4),
                           "how": "rank 0 wall clock over the K back-to-back calls + one device synchronize: the reference bench's "
                                  "definition (cpp/bench/wholememory_ops/gather_scatter_bench.cu:363-366); informational"},
        }
        if e2e:
            line["e2e"] = e2e
        if args.impl == "reference":
            line["impl"] = "reference"
            line["cpu_baseline"] = {"value": round(out_gbs, 3), "unit": "GB/s", "cores": 1, "kind": "reference",
                                    "sample": "reference gather kernels (cpp/src/wholememory_ops) rebuilt for sm_100 from /root/reference "
                                              "into oracle/_ref, same config, same harness; the reference path runs on the GPU, one host thread drives it"}
            line["e2e"] = line.get("e2e") or {"value": round(out_gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        elif not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(dim, esize, np.float32 if esize == 4 else np.float16)
            except Exception as ex:  # the GPU measurement above must not be lost to a host-side problem
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": "unavailable: %r" % (ex,)}
        print(json.dumps(line), flush=True)

    wmb.destroy_wholememory_tensor(table)
    wmb.destroy_communicator(comm)
    if world > 1:
        dist.destroy_process_group()


def reference_cpu_arm(args, rank, world):
    """Fallback reference arm when oracle/_ref is absent: the oracle port of the reference gather on the host cores."""
    if rank != 0:
        return
    import numpy as np
    esize = 4 if args.dtype == "fp32" else 2
    cb = cpu_baseline(args.dim, esize, np.float32 if esize == 4 else np.float16, budget_s=20.0)
    line = {"metric": METRIC, "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(BATCH * args.dim * esize / (cb["value"] * 1e9) * 1e3, 4) if cb["value"] else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if esize == 4 else "f16", "data": "synthetic",
            "impl": "reference", "config": {"workload": "host gather of %d random rows x %d B (oracle port; oracle/_ref not built)" % (BATCH, args.dim * esize)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
