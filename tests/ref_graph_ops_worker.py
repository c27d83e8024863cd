"""Runs seeded graph_append_unique / csr_add_self_loop cases through WHICHEVER build of the C ABI is loaded
(WHOLEGRAPH_B200_LIB = oracle/_ref/libwholegraph_ref.so selects the reference's own graph_ops kernels,
cpp/src/graph_ops/*) and dumps the outputs.  Used by test_zz_ref_graph_ops_parity_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (target count, neighbor count, id dtype)
SMALL = os.environ.get("WG_GOLDEN_SMALL") == "1"  # compact sizes for the committed golden vectors (tools/make_golden.sh)
UNIQUE_CASES = [(10, 100, np.int32), (113, 1987, np.int64), (64, 7, np.int32), (2000, 20000, np.int64) if SMALL else (20000, 200000, np.int64),
                (1, 5000, np.int32)]
# (rows, max degree)
LOOP_CASES = [(1, 5), (37, 9), (5000, 40)]


def unique_inputs(ci):
    t, n, dt = UNIQUE_CASES[ci]
    rng = np.random.default_rng(900 + ci)
    pool = rng.permutation(max(4 * (t + n), 16))
    targets = pool[:t].astype(dt)  # distinct, as the samplers produce them
    neighbors = rng.choice(pool[: max(2 * t, 8) + n // 3 + 1], size=n).astype(dt)  # overlaps targets and repeats
    return targets, neighbors


def loop_inputs(ci):
    rows, max_deg = LOOP_CASES[ci]
    rng = np.random.default_rng(700 + ci)
    deg = rng.integers(0, max_deg + 1, size=rows)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, rows, size=int(row_ptr[-1])).astype(np.int32)
    return row_ptr, col


def run_all(out_path):
    import torch
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.graph_ops import add_csr_self_loop, append_unique
    torch.cuda.set_device(0)
    wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
    out = {}
    for ci in range(len(UNIQUE_CASES)):
        targets, neighbors = unique_inputs(ci)
        tt = torch.from_numpy(targets).cuda() if targets.size else torch.empty(0, dtype=torch.int64 if targets.dtype == np.int64 else torch.int32, device="cuda")
        nt = torch.from_numpy(neighbors).cuda() if neighbors.size else torch.empty(0, dtype=tt.dtype, device="cuda")
        uniq, mapping = append_unique(tt, nt, need_neighbor_raw_to_unique=True)
        torch.cuda.synchronize()
        out["unique_%d" % ci] = uniq.cpu().numpy()
        out["mapping_%d" % ci] = mapping.cpu().numpy()
    for ci in range(len(LOOP_CASES)):
        row_ptr, col = loop_inputs(ci)
        ct = torch.from_numpy(col).cuda() if col.size else torch.empty(0, dtype=torch.int32, device="cuda")
        orow, ocol = add_csr_self_loop(torch.from_numpy(row_ptr).cuda(), ct)
        torch.cuda.synchronize()
        out["loop_row_%d" % ci] = orow.cpu().numpy()
        out["loop_col_%d" % ci] = ocol.cpu().numpy()
    np.savez_compressed(out_path, **out)


if __name__ == "__main__":
    run_all(sys.argv[1])
    print("graph ops worker done:", os.environ.get("WHOLEGRAPH_B200_LIB", "libwholegraph.so (this repo)"))
