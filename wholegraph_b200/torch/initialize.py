"""init / finalize with the public names of pylibwholegraph/torch/initialize.py.

init_torch_env brings up (or reuses) an env:// torch.distributed process group -- nccl for GPU runs, gloo when only the
host control plane is exercised on a CPU box -- which the communicator helpers use to pass unique ids around."""
import os

import torch

from .. import binding as wmb
from .comm import get_global_communicator, get_local_node_communicator, reset_communicators, set_world_info
from .utils import str_to_wmb_wholememory_log_level


def init(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="info"):
    """Initialise the library and record where this process sits in the job (no process group is created)."""
    wmb.init(0, str_to_wmb_wholememory_log_level(wm_log_level))
    set_world_info(world_rank, world_size, local_rank, local_size)


def init_torch_env(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="info",
                   backend: str = "nccl"):
    """init() plus the torch side: rendezvous variables (MASTER_ADDR / MASTER_PORT default to 127.0.0.1:12335), one torch
    thread, the CUDA device of this rank (nccl only) and the process group."""
    os.environ.update({"RANK": str(world_rank), "WORLD_SIZE": str(world_size)})
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12335")
    init(world_rank, world_size, local_rank, local_size, wm_log_level)
    torch.set_num_threads(1)
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend=backend, init_method="env://")


def init_torch_env_and_create_wm_comm(world_rank: int, world_size: int, local_rank: int, local_size: int,
                                      distributed_backend_type="nccl", wm_log_level="info", backend: str = "nccl"):
    """init_torch_env, then (global communicator, this node's communicator)."""
    init_torch_env(world_rank, world_size, local_rank, local_size, wm_log_level, backend)
    return get_global_communicator(distributed_backend_type), get_local_node_communicator()


def finalize():
    """Tear down the library, forget the cached communicators and leave the process group."""
    wmb.finalize()
    reset_communicators()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
