"""init / finalize (mirror of pylibwholegraph/torch/initialize.py)."""
import os

import torch

from .. import binding as wmb
from .comm import get_global_communicator, get_local_node_communicator, reset_communicators, set_world_info
from .utils import str_to_wmb_wholememory_log_level


def init(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="info"):
    wmb.init(0, str_to_wmb_wholememory_log_level(wm_log_level))
    set_world_info(world_rank, world_size, local_rank, local_size)


def init_torch_env(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="info",
                   backend: str = "nccl"):
    """Init WholeGraph for PyTorch: env:// process group (nccl by default, gloo for CPU-only control-plane use)."""
    os.environ["RANK"] = str(world_rank)
    os.environ["WORLD_SIZE"] = str(world_size)
    if "MASTER_ADDR" not in os.environ:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
    if "MASTER_PORT" not in os.environ:
        os.environ["MASTER_PORT"] = "12335"
    wmb.init(0, str_to_wmb_wholememory_log_level(wm_log_level))
    torch.set_num_threads(1)
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend=backend, init_method="env://")
    set_world_info(world_rank, world_size, local_rank, local_size)


def init_torch_env_and_create_wm_comm(world_rank: int, world_size: int, local_rank: int, local_size: int,
                                      distributed_backend_type="nccl", wm_log_level="info", backend: str = "nccl"):
    init_torch_env(world_rank, world_size, local_rank, local_size, wm_log_level, backend)
    global_comm = get_global_communicator(distributed_backend_type)
    local_comm = get_local_node_communicator()
    return global_comm, local_comm


def finalize():
    wmb.finalize()
    reset_communicators()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
