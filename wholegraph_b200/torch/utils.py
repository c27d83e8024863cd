"""dtype / enum string helpers (mirror of pylibwholegraph/torch/utils.py)."""
import torch

from .. import binding as wmb

WholeMemoryDataType = wmb.WholeMemoryDataType

_TORCH_TO_WM = {
    torch.float32: WholeMemoryDataType.DtFloat, torch.float16: WholeMemoryDataType.DtHalf,
    torch.float64: WholeMemoryDataType.DtDouble, torch.bfloat16: WholeMemoryDataType.DtBF16,
    torch.int32: WholeMemoryDataType.DtInt, torch.int64: WholeMemoryDataType.DtInt64,
    torch.int16: WholeMemoryDataType.DtInt16, torch.int8: WholeMemoryDataType.DtInt8,
}
_WM_TO_TORCH = {v: k for k, v in _TORCH_TO_WM.items()}


def torch_dtype_to_wholememory_dtype(torch_dtype):
    if torch_dtype not in _TORCH_TO_WM:
        raise ValueError("torch_dtype: %s not supported" % (torch_dtype,))
    return _TORCH_TO_WM[torch_dtype]


def wholememory_dtype_to_torch_dtype(wm_dtype):
    wm_dtype = WholeMemoryDataType(wm_dtype)
    if wm_dtype not in _WM_TO_TORCH:
        raise ValueError("Invalid wholememory dtype %s" % (wm_dtype,))
    return _WM_TO_TORCH[wm_dtype]


def str_to_wmb_wholememory_memory_type(strmt):
    table = {"continuous": wmb.WholeMemoryMemoryType.MtContinuous, "chunked": wmb.WholeMemoryMemoryType.MtChunked,
             "distributed": wmb.WholeMemoryMemoryType.MtDistributed, "hierarchy": wmb.WholeMemoryMemoryType.MtHierarchy}
    if strmt not in table:
        raise ValueError("WholeMemory type %s not supported, should be (continuous, chunked, distributed, hierarchy)" % strmt)
    return table[strmt]


def str_to_wmb_wholememory_location(str_location):
    table = {"cuda": wmb.WholeMemoryMemoryLocation.MlDevice, "cpu": wmb.WholeMemoryMemoryLocation.MlHost}
    if str_location not in table:
        raise ValueError("WholeMemory location %s not supported, should be (cuda, cpu)" % str_location)
    return table[str_location]


def str_to_wmb_wholememory_log_level(str_log_level):
    table = {"error": wmb.WholeMemoryLogLevel.LevError, "warn": wmb.WholeMemoryLogLevel.LevWarn,
             "info": wmb.WholeMemoryLogLevel.LevInfo, "debug": wmb.WholeMemoryLogLevel.LevDebug,
             "trace": wmb.WholeMemoryLogLevel.LevTrace}
    if str_log_level not in table:
        raise ValueError("WholeMemory log level %s not supported" % str_log_level)
    return table[str_log_level]


def str_to_wmb_wholememory_distributed_backend_type(str_wmb_type):
    table = {"nccl": wmb.WholeMemoryDistributedBackend.DbNCCL, "nvshmem": wmb.WholeMemoryDistributedBackend.DbNVSHMEM}
    if str_wmb_type not in table:
        raise ValueError("WholeMemory backend %s not supported, should be (nccl, nvshmem)" % str_wmb_type)
    return table[str_wmb_type]


def wholememory_distributed_backend_type_to_str(wmb_type):
    names = {int(wmb.WholeMemoryDistributedBackend.DbNCCL): "nccl", int(wmb.WholeMemoryDistributedBackend.DbNVSHMEM): "nvshmem"}
    if int(wmb_type) not in names:
        raise ValueError("WholeMemory distributed backend %s has no name, should be (DbNCCL, DbNVSHMEM)" % (wmb_type,))
    return names[int(wmb_type)]


def str_to_wmb_wholememory_access_type(str_access):
    table = {"readonly": wmb.WholeMemoryAccessType.AtReadOnly, "readwrite": wmb.WholeMemoryAccessType.AtReadWrite}
    if str_access not in table:
        raise ValueError("WholeMemory access type %s not supported" % str_access)
    return table[str_access]


def str_to_wmb_wholememory_optimizer_type(str_opt):
    table = {"sgd": wmb.WholeMemoryOptimizerType.OptSgd, "adam": wmb.WholeMemoryOptimizerType.OptLazyAdam,
             "adagrad": wmb.WholeMemoryOptimizerType.OptAdaGrad, "rmsprop": wmb.WholeMemoryOptimizerType.OptRmsProp}
    if str_opt not in table:
        raise ValueError("WholeMemory optimizer type %s not supported, should be (sgd, adam, adagrad, rmsprop)" % str_opt)
    return table[str_opt]


def get_file_size(filename: str):
    import os
    if not os.path.isfile(filename):
        raise ValueError("File %s not found or not file" % (filename,))
    if not os.access(filename, os.R_OK):
        raise ValueError("File %s not readable" % (filename,))
    return os.path.getsize(filename)


def get_part_file_name(prefix: str, part_id: int, part_count: int):
    return "%s_part_%d_of_%d" % (prefix, part_id, part_count)


def get_part_file_list(prefix: str, part_count: int):
    return [get_part_file_name(prefix, part_id, part_count) for part_id in range(part_count)]


def count_file_entries(filelist, entry_bytes: int) -> int:
    """Entries of `entry_bytes` bytes held by a list of raw binary files; every file must hold a whole number of them
    (the rule both file-backed constructors of the reference apply, tensor.py:276-287 / embedding.py:498-509)."""
    total = 0
    for filename in filelist:
        size = get_file_size(filename)
        if size % entry_bytes != 0:
            raise ValueError("File %s size is %d not mutlple of %d" % (filename, size, entry_bytes))
        total += size
    return total // entry_bytes
