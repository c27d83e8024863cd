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
