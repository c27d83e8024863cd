"""Runs seeded unweighted neighbor-sampling cases through WHICHEVER build of the C ABI is loaded and dumps all four
outputs.  With WHOLEGRAPH_B200_LIB = oracle/_ref/libwholegraph_ref.so these are the reference's own kernels
(cpp/src/wholegraph_ops/unweighted_sample_without_replacement_func.cuh: count -> scan -> BlockRadixSort + pointer-jumping
sampler, sample_comm.cuh sample-all) compiled on the RESTATED PCG stand-in (oracle/ref_shim/raft/random/rng_device.cuh).
Used by test_zz_ref_sampling_parity_gpu.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
sys.path.insert(0, os.path.join(ROOT, "tests"))

NODES = 4000
# (max_sample_count, col dtype, center dtype, centers, seed): k around the reference's launch-shape boundaries
# ((k-1)/32 picks BLOCK_DIM / ITEMS_PER_THREAD, func.cuh:423-458), k <= 0 = take all neighbours
CASES = [
    (5, np.int32, np.int32, 700, 11),
    (10, np.int64, np.int64, 1024, 12),
    (25, np.int32, np.int32, 1024, 13),
    (32, np.int64, np.int32, 513, 14),
    (33, np.int32, np.int64, 300, 15),
    (64, np.int64, np.int64, 300, 16),
    (100, np.int32, np.int32, 200, 17),
    (257, np.int64, np.int64, 100, 18),
    (-1, np.int32, np.int32, 400, 19),
]


def graph(col_dtype):
    """Degrees 0 .. ~600 (heavy tail) so every case mixes deg <= k (copy) and deg > k (sample) centers."""
    rng = np.random.default_rng(2024)
    deg = np.minimum((rng.pareto(1.1, size=NODES) * 8).astype(np.int64), 600)
    deg[::17] = 0
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, NODES, size=int(row_ptr[-1])).astype(col_dtype)
    return row_ptr, col


def centers_of(ci):
    k, cdt, ndt, n, seed = CASES[ci]
    return np.random.default_rng(500 + ci).integers(0, NODES, size=n).astype(ndt)


def run_all(out_path):
    import torch
    import gpu_utils as G
    import wholegraph_b200.torch as wgth
    from wholegraph_b200.torch.wholegraph_ops import unweighted_sample_without_replacement
    comm = wgth.WholeMemoryCommunicator(G.single_comm())
    out = {}
    tensors = {}
    for cdt in (np.int32, np.int64):
        row_ptr, col = graph(cdt)
        rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [NODES + 1], torch.int64, [1])
        cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [col.size], torch.int32 if cdt == np.int32 else torch.int64, [1])
        rp.get_local_tensor()[0].copy_(torch.from_numpy(row_ptr))
        cp.get_local_tensor()[0].copy_(torch.from_numpy(col))
        tensors[cdt] = (rp, cp)
    torch.cuda.synchronize()
    for ci, (k, cdt, ndt, n, seed) in enumerate(CASES):
        rp, cp = tensors[cdt]
        res = unweighted_sample_without_replacement(rp.wmb_tensor, cp.wmb_tensor, torch.from_numpy(centers_of(ci)).cuda(), k,
                                                    random_seed=seed, need_center_local_output=True, need_edge_output=True)
        torch.cuda.synchronize()
        for name, t in zip(("offsets", "dst", "center_lid", "edge_gid"), res):
            out["case%d_%s" % (ci, name)] = t.cpu().numpy()
    for rp, cp in tensors.values():
        wgth.destroy_wholememory_tensor(rp)
        wgth.destroy_wholememory_tensor(cp)
    np.savez_compressed(out_path, **out)


if __name__ == "__main__":
    run_all(sys.argv[1])
    print("sampling worker done:", os.environ.get("WHOLEGRAPH_B200_LIB", "libwholegraph.so (this repo)"))
