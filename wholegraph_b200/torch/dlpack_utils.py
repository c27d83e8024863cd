"""DLPack import helper (mirror of pylibwholegraph/torch/dlpack_utils.py)."""
import torch.utils.dlpack


def torch_import_from_dlpack(dp):
    return torch.utils.dlpack.from_dlpack(dp.__dlpack__()) if hasattr(dp, "__dlpack__") else torch.utils.dlpack.from_dlpack(dp)
