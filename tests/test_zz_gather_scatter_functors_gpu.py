"""The reference's torch-level gather / scatter op test (test_wholegraph_gather_scatter.py) with its own sizes, through
wholegraph_b200.torch.wholememory_ops and the reference's `get_local_tensor(import_dlpack_fn, location, device)` call form.
One rank and two ranks sharing the GPU.

(File name sorts last on purpose: added without a GPU at hand; the verified gather / scatter tests are
tests/test_gather_scatter_gpu.py and the `gather_scatter` scenario of tests/test_multi_rank_gpu.py.)"""
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [1, 2])
def test_reference_torch_level_gather_scatter(world):
    import test_multi_rank_gpu as M
    M._run(world, "gather_scatter_functors", share_gpu=world > 1)
