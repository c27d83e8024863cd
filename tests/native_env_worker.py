"""GPU half of test_zz_native_env.py, run in its own process (an allocator-callback bug would be a crash, not an
assertion): sampling + append_unique through the Python-callback env table, then three times through the native (C++)
table -- same results, no leaked contexts."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import gpu_utils as G
    import wholegraph_b200.torch as wgth
    from wholegraph_b200.torch import wholegraph_env as wenv
    from wholegraph_b200.torch.wholegraph_ops import unweighted_sample_without_replacement

    comm = wgth.WholeMemoryCommunicator(G.single_comm())
    rng = np.random.default_rng(3)
    nodes = 3000
    deg = rng.integers(0, 60, size=nodes)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, nodes, size=int(row_ptr[-1])).astype(np.int64)
    rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
    cp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [col.size], torch.int64, [1])
    rp.get_local_tensor()[0].copy_(torch.from_numpy(row_ptr))
    cp.get_local_tensor()[0].copy_(torch.from_numpy(col))
    centers = torch.from_numpy(rng.integers(0, nodes, size=2048).astype(np.int64)).cuda()
    targets = torch.from_numpy(rng.permutation(nodes)[:500].astype(np.int64)).cuda()

    def one_pass():
        res = unweighted_sample_without_replacement(rp.wmb_tensor, cp.wmb_tensor, centers, 25, random_seed=99,
                                                    need_center_local_output=True, need_edge_output=True)
        uniq, mapping = wgth.append_unique(targets, res[1], need_neighbor_raw_to_unique=True)
        torch.cuda.synchronize()
        return [r.cpu() for r in res] + [uniq.cpu(), mapping.cpu()]

    wenv.unload_native_env()
    expect = one_pass()
    assert wenv.load_native_env(required=True)
    lib = wenv.torch_cpp_ext_lib
    before = lib.live_context_count()
    assert lib.get_stream() == wenv.get_stream()
    for _ in range(3):
        got = one_pass()
        assert len(got) == len(expect) and all(torch.equal(a, b) for a, b in zip(got, expect))
    assert lib.live_context_count() == before
    wenv.unload_native_env()
    wgth.destroy_wholememory_tensor(rp)
    wgth.destroy_wholememory_tensor(cp)
    print("native env worker OK")


if __name__ == "__main__":
    main()
