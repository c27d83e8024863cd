"""The division-free "vector index -> (row, vector in row)" map of the row-move kernels, checked exhaustively at the row
boundaries on CPU.

Host side (wholegraph_b200/csrc/gather_scatter.cu: set_units): magic = ceil(2^40 / units_per_row).  Only batches of
more than one row use the map, and plan() batches rows only while batch_rows * row_bytes <= 4 KiB, so the units_per_row
that reach it are < 2^12; the test covers every units_per_row * 32 < 2^20 (the range round 1 accepted).  Longer rows travel
one per warp (row = 0, no map; tests/test_access_width_gpu.py::test_very_long_rows).  Device side (gather_scatter.cuh: row_move_vec_kernel / row_move_cvt_kernel):
row = (w * magic) >> 40 for w in [0, batch_rows * units_per_row) plus up to 32 * UNROLL lanes of overshoot.
The claim in table_ref.hpp is that this equals floor(w / units_per_row) on that whole range."""
import numpy as np


def test_magic_division_is_exact_for_every_accepted_row_length():
    d = np.arange(1, (1 << 20) // 32, dtype=np.uint64)                  # every units_per_row set_units accepts
    magic = ((np.uint64(1) << np.uint64(40)) + d - np.uint64(1)) // d
    overshoot = 32 * 8                                                  # lanes past the end still compute a row number
    for r in range(0, 34):                                              # batch_rows <= 32: rows 0..32, and one beyond
        for delta in (-1, 0, 1, overshoot):
            w = d * np.uint64(r) + np.uint64(delta) if delta >= 0 else d * np.uint64(r) - np.uint64(1)
            if r == 0 and delta < 0:
                continue
            assert int(w.max()) < (1 << 24)                             # w * magic stays far below 2^64
            got = (w * magic) >> np.uint64(40)
            assert np.array_equal(got, w // d), (r, delta)


def test_magic_division_random_interior_points():
    rng = np.random.default_rng(1)
    d = rng.integers(1, (1 << 20) // 32, size=2_000_000).astype(np.uint64)
    w = (rng.random(2_000_000) * (d.astype(np.float64) * 32 + 256)).astype(np.uint64)
    magic = ((np.uint64(1) << np.uint64(40)) + d - np.uint64(1)) // d
    assert np.array_equal((w * magic) >> np.uint64(40), w // d)


def test_rotated_owner_walk_of_the_gradient_push_is_a_permutation():
    """peer_push.cu: warp j of the push grid handles grouped position (j + bucket_start[(me + 1) % ws]) mod n_send.  For any
    bucket sizes (empty buckets included) that is a bijection on [0, n_send), it starts at owner me + 1's bucket (or the
    next non-empty one) and ends with the rank's own bucket."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        ws = int(rng.integers(1, 17))
        me = int(rng.integers(0, ws))
        counts = rng.integers(0, 50, size=ws) * (rng.random(ws) > 0.3)
        start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        n = int(start[-1])
        if n == 0:
            continue
        first = int(start[(me + 1) % ws])
        j = np.arange(n, dtype=np.int64)
        pos = j + first
        pos = np.where(pos >= n, pos - n, pos)
        assert sorted(pos.tolist()) == list(range(n))
        owner = np.searchsorted(start, pos, side="right") - 1
        order = [int(o) for i, o in enumerate(owner) if i == 0 or owner[i] != owner[i - 1]]
        expect = [r % ws for r in range(me + 1, me + 1 + ws) if counts[r % ws] > 0]
        assert order == expect, (order, expect)


def test_long_run_dispatch_lists_every_hot_row_exactly_once():
    """sparse_optimizer.cu: the head of a run (sorted position b with sorted[b-1] != sorted[b]) hands its row to
    long_run_update_kernel iff b + 64 < n and sorted[b + 64] == sorted[b], i.e. iff the run is longer than 64; the run's
    end is then the upper bound found by the binary search that starts at b + 64.  Checked against run lengths from numpy."""
    rng = np.random.default_rng(4)
    for _ in range(200):
        lens = rng.choice([1, 2, 63, 64, 65, 66, 127, 128, 129, 700], size=int(rng.integers(1, 30)))
        ids = np.sort(rng.choice(10**6, size=lens.size, replace=False))
        s = np.repeat(ids, lens)
        n = s.size
        heads = [b for b in range(n) if b == 0 or s[b - 1] != s[b]]
        listed = [b for b in heads if b + 64 < n and s[b + 64] == s[b]]
        truth = [int(h) for h, L in zip(np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) if L > 64]
        assert listed == truth
        for b in listed:
            lo, hi = b + 64, n
            while hi - lo > 1:
                mid = lo + (hi - lo) // 2
                if s[mid] == s[b]:
                    lo = mid
                else:
                    hi = mid
            assert hi - b == lens[heads.index(b)]
