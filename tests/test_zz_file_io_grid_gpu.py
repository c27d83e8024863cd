"""The reference's part-file load / store test grid (test_wholememory_io.py) through this repo's torch layer: column windows of
a wider table (storage offset 3), dims 16 / 31 / 33 in strides 32 / 64, 3 / 5 part files, default and random row partitions,
CONTINUOUS / CHUNKED / DISTRIBUTED, device and host memory.  One rank, and 2 / 3 ranks sharing the GPU.

(File name sorts last on purpose: added without a GPU at hand; the verified round-trip test is the `file_io` scenario of
tests/test_multi_rank_gpu.py.)"""
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_part_file_load_store_grid(world):
    import test_multi_rank_gpu as M
    M._run(world, "file_io_grid", share_gpu=world > 1)
