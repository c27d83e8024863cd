"""The reference's Python sampling test grids (unweighted and weighted) through this repo's torch layer: CSR in DEVICE and HOST
memory, CONTINUOUS / CHUNKED / DISTRIBUTED, int32 / int64 ids, float / double weights, fan-out 11 and -1, and all four
combinations of the optional outputs, checked against the oracle.  One rank and two ranks sharing the GPU.

(File name sorts last on purpose: added without a GPU at hand; the verified sampler tests are the `sampling` and
`weighted_sampling` scenarios of tests/test_multi_rank_gpu.py.)"""
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [1, 2])
def test_reference_sampling_grids(world):
    import test_multi_rank_gpu as M
    M._run(world, "sampling_grid", share_gpu=world > 1)
