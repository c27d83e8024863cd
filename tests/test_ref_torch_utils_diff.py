"""pylibwholegraph/torch/utils.py of the reference (loaded unchanged, its binding import resolved to this repo's ctypes
binding through compat/) against wholegraph_b200/torch/utils.py: every string -> enum map, the dtype maps in both
directions, the part-file naming and the file-size helper, over valid and invalid inputs.  CPU only."""
import os

import pytest
import torch

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/utils.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


def _both():
    from compat_loader import load_reference_file
    import wholegraph_b200.torch.utils as ours
    return ours, load_reference_file(REF, "_reference_torch_utils")


def _outcome(fn, *args):
    try:
        return ("ok", fn(*args))
    except Exception as e:  # the kind of failure is part of the contract, the message is not
        return ("raises", type(e).__name__)


def test_string_and_dtype_maps_agree(tmp_path):
    ours, ref = _both()
    import wholegraph_b200.binding as wmb
    words = ["continuous", "chunked", "distributed", "hierarchy", "cpu", "cuda", "host", "device", "readonly", "readwrite", "sgd", "adam",
             "adagrad", "rmsprop", "lazy_adam", "nccl", "nvshmem", "none", "debug", "info", "warn", "error", "fatal", "trace", "", "CUDA", "Adam", 3]
    for name in ("str_to_wmb_wholememory_memory_type", "str_to_wmb_wholememory_location", "str_to_wmb_wholememory_log_level",
                 "str_to_wmb_wholememory_access_type", "str_to_wmb_wholememory_optimizer_type",
                 "str_to_wmb_wholememory_distributed_backend_type"):
        for w in words:
            assert _outcome(getattr(ours, name), w) == _outcome(getattr(ref, name), w), (name, w)
    for dt in (torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8,
               torch.bool, torch.complex64):
        assert _outcome(ours.torch_dtype_to_wholememory_dtype, dt) == _outcome(ref.torch_dtype_to_wholememory_dtype, dt), dt
    for wd in list(wmb.WholeMemoryDataType):
        assert _outcome(ours.wholememory_dtype_to_torch_dtype, wd) == _outcome(ref.wholememory_dtype_to_torch_dtype, wd), wd
    for be in list(wmb.WholeMemoryDistributedBackend):
        assert _outcome(ours.wholememory_distributed_backend_type_to_str, be) == _outcome(ref.wholememory_distributed_backend_type_to_str, be), be
    for prefix, parts in (("emb", 1), ("/data/x_y", 8), ("", 3)):
        assert ours.get_part_file_list(prefix, parts) == ref.get_part_file_list(prefix, parts)
        for i in range(parts):
            assert ours.get_part_file_name(prefix, i, parts) == ref.get_part_file_name(prefix, i, parts)
    f = tmp_path / "blob.bin"
    f.write_bytes(b"\0" * 12345)
    assert ours.get_file_size(str(f)) == ref.get_file_size(str(f)) == 12345
    assert _outcome(ours.get_file_size, str(tmp_path / "missing"))[0] == _outcome(ref.get_file_size, str(tmp_path / "missing"))[0]
