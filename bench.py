#!/usr/bin/env python
"""bench.py -- device-timed WholeMemory embedding gather (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                      # N=1: configs[1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               # the reference's own kernels (oracle/_ref)

A "step" is ONE wholememory_gather call over one batch of 1,048,576 uniform-random int64 indices.

N=1 workload = BASELINE.json configs[1]: CONTINUOUS/DEVICE table 100M x 256 fp32 (102.4 GB) on one B200.
N>1 (weak scaling): CHUNKED/DEVICE table of N x 100M rows x 256 fp32 row-sharded over the N GPUs of the box
and mapped into every GPU (VMM over NVSwitch); every rank gathers 1M rows drawn uniformly from the WHOLE table.
After the headline, the same measurement runs on the other BASELINE shapes at their per-GPU size (device-timed only):
  ns = north star, 125M rows/GPU x 256 fp16 (1B x 256 fp16 at N=8);  c3 = configs[2], 125M rows/GPU x 128 fp16.
Their results are listed under config.other_shapes (and top-level "shapes"); --shapes c2 runs the headline alone.

value       = n * row_out_bytes * N / t   (GB/s of gathered output -- the reference bench's "Bandwidth",
              cpp/bench/wholememory_ops/gather_scatter_bench.cu:363-366), t = max over ranks of the CUDA-event
              time of the K timed calls / K, inputs resident in HBM.
e2e         = same metric with the step's indices starting in pinned HOST memory (H2D inside the timed region)
              and the gathered rows copied back to pinned host memory (D2H inside the timed region).
roofline    = N=1: algorithmic bytes n*(row_in + row_out + idx) / kernel time against the measured HBM copy bandwidth.
              N>1: the bound is NVLink ingress -- remote bytes n*row*(N-1)/N per GPU / kernel time against the measured
              peer-copy bandwidth (770 GB/s per direction, B200_PROFILING.md; 900 nominal also given).
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
    p.add_argument("--shapes", default="c2,ns,c3", help="comma list of c2 (headline), ns, c3; the first one is the headline line")
    p.add_argument("--rows-per-gpu", type=int, default=None, help="custom single shape (diagnostics): overrides --shapes")
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


SHAPES = {
    # BASELINE.json configs[1] (weak-scaled for N > 1): the headline
    "c2": {"rows_per_gpu": 100_000_000, "dim": 256, "dtype": "fp32", "name": "C2: 100M rows/GPU x 256 fp32 (1 KiB rows)"},
    # north_star: 1B x 256 fp16 over 8 GPUs = 125M rows per GPU
    "ns": {"rows_per_gpu": 125_000_000, "dim": 256, "dtype": "fp16", "name": "north star: 125M rows/GPU x 256 fp16 (512 B rows; 1B rows at N=8)"},
    # BASELINE.json configs[2]: 1B x 128 fp16 over 8 GPUs
    "c3": {"rows_per_gpu": 125_000_000, "dim": 128, "dtype": "fp16", "name": "C3: 125M rows/GPU x 128 fp16 (256 B rows; 1B rows at N=8)"},
}
NVLINK_MEASURED_GBS = 770.0  # peer copy per direction measured on this pool (B200_PROFILING.md); nominal 900
# ncu NVLink counters of this kernel (profiles/r2_nvlink_counters.txt): a peer load puts 1.125 bytes per payload byte on the
# reader's RX side (16 B response header per 128 B) and 0.1875 on its TX side (24 B request per 128 B); one-directional reads
# fill the link at 880 GB/s raw = 783 GB/s of payload.  When every GPU reads from every other GPU, each direction carries
# both: payload <= 880 / 1.3125 = 670.6 GB/s per direction -- the ceiling of the all-to-all gather, whatever the kernel does.
NVLINK_RAW_GBS = 880.0
NVLINK_WIRE_BYTES_PER_PAYLOAD_BYTE_ALL_TO_ALL_READS = 1.3125


def traffic_from_profile(shape_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of THIS bench command
    on the bench-sized table (profiles/r2_gather_<shape>_full_summary.txt, written by tools/ncu_summary.py)."""
    path = os.path.join(ROOT, "profiles", "r2_gather_%s_full_summary.txt" % shape_key)
    if not os.path.exists(path):
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    got = {}
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[0] not in got and f[2] in unit:
            got[f[0]] = float(f[1].replace(",", "")) * unit[f[2]]
    if len(got) != 2:
        return None, None
    return got["dram__bytes_read.sum"] + got["dram__bytes_write.sum"], \
        "profiles/r2_gather_%s_full_summary.txt (ncu --set full of this command: %.3f GB read + %.3f GB written per launch)" % (
            shape_key, got["dram__bytes_read.sum"] / 1e9, got["dram__bytes_write.sum"] / 1e9)


def roofline(world, n, row, ms, impl, shape_key):
    hbm_peak, peak_src = measured_peaks()
    alg_bytes = n * (row + row + 8)
    alg_gbs = alg_bytes / (ms * 1e-3) / 1e9  # per GPU
    kernel = "wm::row_move_vec_kernel<int64, 32 B units, gather> (wholegraph_b200/csrc/gather_scatter.cuh)" if impl != "reference" else \
        "reference gather_func_kernel / gather_func_sub_warp_kernel (cpp/src/wholememory_ops/functions/gather_scatter_func.cuh, rebuilt in oracle/_ref)"
    if world == 1:
        traffic, traffic_src = traffic_from_profile(shape_key) if impl != "reference" and n == BATCH else (None, None)
        return {"bound": "hbm", "achieved": round(alg_gbs, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(alg_gbs / hbm_peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel, "kernel_ms": round(ms, 4),
                "algorithmic_bytes_per_launch": alg_bytes}
    remote_bytes = n * row * (world - 1) / world  # uniform indices over an equally partitioned table
    ingress = remote_bytes / (ms * 1e-3) / 1e9
    t_nvl, t_hbm = remote_bytes / (NVLINK_MEASURED_GBS * 1e9), alg_bytes / (hbm_peak * 1e9)
    return {"bound": "nvlink", "achieved": round(ingress, 2), "peak": NVLINK_MEASURED_GBS, "unit": "GB/s", "frac": round(ingress / NVLINK_MEASURED_GBS, 4),
            "frac_of_nominal_900": round(ingress / 900.0, 4), "traffic": None, "traffic_source": None,
            "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md: 770 GB/s; nominal NVLink 5 = 900 GB/s)",
            "what": "NVLink ingress per GPU = the (N-1)/N of the gathered rows that live on peers / kernel time",
            "protocol_ceiling": round(NVLINK_RAW_GBS / NVLINK_WIRE_BYTES_PER_PAYLOAD_BYTE_ALL_TO_ALL_READS, 1),
            "frac_of_protocol_ceiling": round(ingress / (NVLINK_RAW_GBS / NVLINK_WIRE_BYTES_PER_PAYLOAD_BYTE_ALL_TO_ALL_READS), 4),
            "protocol_ceiling_source": "profiles/r2_nvlink_counters.txt: 880 GB/s raw link rate / 1.3125 wire bytes per payload byte when all GPUs read at once",
            "kernel": kernel, "kernel_ms": round(ms, 4), "bound_ms": round(max(t_nvl, t_hbm) * 1e3, 4),
            "hbm_algorithmic_gbs": round(alg_gbs, 2), "hbm_frac": round(alg_gbs / hbm_peak, 4), "remote_bytes_per_launch": int(remote_bytes),
            "algorithmic_bytes_per_launch": alg_bytes}


def run_shape(ctx, key, shape, args, headline):
    """Allocate the table of one shape, time K gathers on the device, check the gathered rows, free the table."""
    import torch
    import torch.distributed as dist
    wmb, env, comm, rank, world, local_rank = ctx["wmb"], ctx["env"], ctx["comm"], ctx["rank"], ctx["world"], ctx["local_rank"]
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    th_dtype, wm_dtype, esize = (torch.float32, wmb.DtFloat, 4) if shape["dtype"] == "fp32" else (torch.float16, wmb.DtHalf, 2)
    dim, n, rows_per_gpu = shape["dim"], args.batch, shape["rows_per_gpu"]
    rows_total = rows_per_gpu * world
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
    del local
    torch.cuda.synchronize()
    comm.barrier()

    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED + rank)
    n_batches = 8  # distinct index batches, cycled: successive steps never re-read the same rows
    idx_dev = [torch.randint(0, rows_total, (n,), device="cuda", dtype=torch.int64, generator=gen) for _ in range(n_batches)]
    if args.index_pattern in ("remote", "local") and world > 1:  # diagnostic: only peer rows / only my rows
        per = rows_per_gpu
        idx_dev = []
        for _ in range(n_batches):
            r = torch.randint(0, per * (world - 1) if args.index_pattern == "remote" else per, (n,), device="cuda", dtype=torch.int64, generator=gen)
            if args.index_pattern == "remote":
                r = r + (r >= rank * per).to(torch.int64) * per  # skip my own partition
            else:
                r = r + rank * per
            idx_dev.append(r)
    if args.index_pattern == "sequential":  # diagnostic only: contiguous rows
        idx_dev = [(torch.arange(n, device="cuda", dtype=torch.int64) + (b * n * 7) % max(1, rows_total - n)) for b in range(n_batches)]
    out = torch.empty(n, dim, device="cuda", dtype=th_dtype)
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

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(i)
    # correctness of the timed configuration (cheap, outside the timed region)
    torch.cuda.synchronize()
    last = (warm - 1) % n_batches
    assert torch.equal(out[:, 0], (idx_dev[last] & mask).to(th_dtype)) and torch.equal(out[:, dim - 1], out[:, 0]), "gather wrong (%s)" % key

    sampler = ClockSampler(local_rank) if (rank == 0 and headline) else None
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
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    clocks = None
    if headline:
        # keep the GPUs busy ~1 s longer so the clock sampler sees the loaded state even for tiny K.  EVERY rank runs the SAME
        # number of extra steps (derived from the max-over-ranks time): on DISTRIBUTED memory without peer mapping -- and always
        # with the reference library -- the gather is a collective call, and a rank issuing more of them than its peers hangs
        extra = max(1, min(5000, int(1000.0 / max(ms_max, 1e-3))))
        for k in range(extra):
            step(k)
        torch.cuda.synchronize()
        if sampler:
            clocks = sampler.summary()
        sync_all()

    if args.per_step_events and rank == 0 and mem_type != "distributed":
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i, (a, b) in enumerate(evs):
            a.record(stream)
            step(i)
            b.record(stream)
        torch.cuda.synchronize()
        per = sorted(a.elapsed_time(b) for a, b in evs)
        print("%s per-step kernel ms: min %.4f median %.4f max %.4f (loop mean %.4f)" % (key, per[0], per[len(per) // 2], per[-1], ms), file=sys.stderr)

    # ---- end to end (headline only): indices from pinned host, rows back to pinned host, both copies inside the timed region
    e2e = None
    if headline and not args.no_e2e:
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
        del out_host, idx_host

    row = dim * esize
    rec = {"key": key, "name": shape["name"], "rows_per_gpu": rows_per_gpu, "rows_total": rows_total, "dim": dim, "dtype": shape["dtype"], "row_bytes": row,
           "memory_type": mem_type, "ms_per_step": round(ms_max, 4), "value": round(n * row * world / (ms_max * 1e-3) / 1e9, 3), "unit": "GB/s",
           "per_gpu_gbs_out": round(n * row / (ms_max * 1e-3) / 1e9, 3), "host_ms": host_ms, "clocks": clocks, "e2e": e2e,
           "roofline": roofline(world, n, row, ms_max, args.impl, key), "checked": "every gathered row equals the closed-form pattern of its index"}
    del w_idx, w_out, out, idx_dev
    comm.barrier()
    wmb.destroy_wholememory_tensor(table)
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # the reference library is loaded THROUGH THE SAME BINDING (identical C ABI): oracle/_ref build
        ref = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")
        if os.path.exists(ref):
            from oracle.ref_lib_loader import use_library
            use_library(ref)
        else:
            return reference_cpu_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_wholegraph_env_fns

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
    ctx = {"wmb": wmb, "env": get_wholegraph_env_fns(), "comm": comm, "rank": rank, "world": world, "local_rank": local_rank}

    if args.rows_per_gpu is not None:  # diagnostics: one custom shape
        plan = [("custom", {"rows_per_gpu": args.rows_per_gpu, "dim": args.dim, "dtype": args.dtype,
                            "name": "custom: %d rows/GPU x %d %s" % (args.rows_per_gpu, args.dim, args.dtype)})]
    else:
        plan = [(k, SHAPES[k]) for k in args.shapes.split(",") if k]
    recs = []
    for i, (key, shape) in enumerate(plan):
        recs.append(run_shape(ctx, key, shape, args, headline=(i == 0)))
    head, n = recs[0], args.batch

    if rank == 0:
        row = head["row_bytes"]
        others = [{k: r[k] for k in ("key", "name", "rows_total", "dim", "dtype", "row_bytes", "memory_type", "value", "unit", "per_gpu_gbs_out",
                                     "ms_per_step", "roofline", "checked")} for r in recs[1:]]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if head["dtype"] == "fp32" else "f16", "data": "synthetic",
            "config": {"workload": "%s/DEVICE gather, %d x %d %s table (%.1f GB per GPU), %d uniform-random int64 indices per rank per step"
                                   % (head["memory_type"].upper(), head["rows_total"], head["dim"], head["dtype"], head["rows_per_gpu"] * row / 1e9, n),
                       "l2": "inputs larger than L2: table %.0f GB, 8 distinct index batches cycled, %.2f GB output per step" % (head["rows_total"] * row / 1e9, n * row / 1e9),
                       "rows_per_gpu": head["rows_per_gpu"], "embedding_dim": head["dim"], "indices_per_rank": n, "memory_type": head["memory_type"],
                       "other_shapes": others},
            "roofline": head["roofline"], "gpu_launches": args.steps, "clocks": head["clocks"],
            "host_timed": {"value": round(n * row * world / (head["host_ms"] * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(head["host_ms"], 4),
                           "how": "rank 0 wall clock over the K back-to-back calls + one device synchronize: the reference bench's "
                                  "definition (cpp/bench/wholememory_ops/gather_scatter_bench.cu:363-366); informational"},
            "shapes": [{"key": r["key"], "value": r["value"], "ms_per_step": r["ms_per_step"], "frac": r["roofline"]["frac"],
                        "bound": r["roofline"]["bound"]} for r in recs],
        }
        if head["e2e"]:
            line["e2e"] = head["e2e"]
        if args.impl == "reference":
            line["impl"] = "reference"
            line["cpu_baseline"] = {"value": head["value"], "unit": "GB/s", "cores": 1, "kind": "reference",
                                    "sample": "reference gather kernels (cpp/src/wholememory_ops) rebuilt for sm_100 from /root/reference "
                                              "into oracle/_ref, same config, same harness; the reference path runs on the GPU, one host thread drives it"}
            line["e2e"] = line.get("e2e") or {"value": head["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        elif not args.no_cpu_baseline and world == 1:
            esize = 4 if head["dtype"] == "fp32" else 2
            try:
                line["cpu_baseline"] = cpu_baseline(head["dim"], esize, np.float32 if esize == 4 else np.float16)
            except Exception as ex:  # the GPU measurement above must not be lost to a host-side problem
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": "unavailable: %r" % (ex,)}
        print(json.dumps(line), flush=True)

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
