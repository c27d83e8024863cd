"""Runs a fixed, seeded list of gather/scatter cases through WHICHEVER build of the C ABI is loaded
(WHOLEGRAPH_B200_LIB selects the reference's own library rebuilt under oracle/_ref) and dumps the raw output
bytes.  Used by test_ref_parity_gpu.py and by tools/make_golden.py (golden fixtures from the reference itself)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (mem_type, location, table dtype, dense dtype, cols, stride, index dtype, n)
from oracle import oracle as O  # noqa: E402  (dtype ids only; the worker never calls the oracle)

CASES = [
    ("continuous", "cuda", O.DT_FLOAT, O.DT_FLOAT, 256, 256, np.int64, 4099),
    ("chunked", "cuda", O.DT_HALF, O.DT_HALF, 128, 128, np.int32, 4099),
    ("chunked", "cuda", O.DT_HALF, O.DT_FLOAT, 127, 127, np.int64, 1000),
    ("continuous", "cuda", O.DT_FLOAT, O.DT_HALF, 129, 132, np.int64, 1005),
    ("continuous", "cuda", O.DT_DOUBLE, O.DT_HALF, 11, 12, np.int32, 1005),
    ("chunked", "cuda", O.DT_DOUBLE, O.DT_FLOAT, 32, 33, np.int64, 777),
    ("continuous", "cuda", O.DT_FLOAT, O.DT_DOUBLE, 1, 1, np.int64, 513),
    ("chunked", "cuda", O.DT_INT64, O.DT_INT8, 13, 16, np.int64, 600),
    ("continuous", "cuda", O.DT_INT8, O.DT_INT, 513, 520, np.int32, 300),
    ("continuous", "cuda", O.DT_INT16, O.DT_INT16, 7, 7, np.int64, 2000),
    ("distributed", "cuda", O.DT_FLOAT, O.DT_FLOAT, 64, 64, np.int64, 3000),
    ("continuous", "cpu", O.DT_FLOAT, O.DT_FLOAT, 64, 64, np.int64, 3000),   # reference: sorted-ids branch (row <= 512 B)
    ("chunked", "cpu", O.DT_HALF, O.DT_FLOAT, 300, 304, np.int32, 1200),
]
ROWS = 6007
if os.environ.get("WG_GOLDEN_SMALL"):  # tools/make_golden.sh: a compact version of the same cases, small enough to commit
    ROWS = 307
    CASES = [(mt, loc, a, b, min(cols, 40), min(stride, 40) if stride != cols else min(cols, 40), idt, min(n, 150))
             for (mt, loc, a, b, cols, stride, idt, n) in CASES]


def case_inputs(ci):
    """Everything random about case ci, reproducible in any process."""
    import gpu_utils as G
    mt, loc, tab_dt, out_dt, cols, stride, idt, n = CASES[ci]
    rng = np.random.default_rng(31337 + ci)
    table = G.random_table(rng, tab_dt, ROWS, stride)
    idx = rng.integers(0, ROWS, size=n).astype(idt)
    sentinel = G.random_table(rng, out_dt, n, cols)
    sidx = rng.permutation(ROWS)[: n // 2].astype(idt)
    src = G.random_table(rng, out_dt, n // 2, cols)
    return table, idx, sentinel, sidx, src


def run_all(out_path):
    import torch
    import gpu_utils as G
    comm = G.single_comm()
    results = {}
    for ci, (mt, loc, tab_dt, out_dt, cols, stride, idt, n) in enumerate(CASES):
        table, idx, sentinel, sidx, src = case_inputs(ci)
        t, view = G.create_table(comm, mt, loc, tab_dt, ROWS, cols, stride)
        view.copy_(G.np_to_torch(table, tab_dt))
        torch.cuda.synchronize()
        out_t = G.np_to_torch(sentinel.copy(), out_dt).cuda()
        G.gather(t, torch.from_numpy(idx).cuda(), out_t)
        torch.cuda.synchronize()
        results["gather_%d" % ci] = np.frombuffer(G.torch_to_np(out_t, out_dt).tobytes(), dtype=np.uint8)
        G.scatter(G.np_to_torch(src, out_dt).cuda(), torch.from_numpy(sidx).cuda(), t)
        torch.cuda.synchronize()
        results["scatter_%d" % ci] = np.frombuffer(G.torch_to_np(view, tab_dt).tobytes(), dtype=np.uint8)
        G.wmb.destroy_wholememory_tensor(t)
    np.savez_compressed(out_path, **results)


if __name__ == "__main__":
    run_all(sys.argv[1])
    print("worker done:", os.environ.get("WHOLEGRAPH_B200_LIB", "libwholegraph.so (this repo)"))
