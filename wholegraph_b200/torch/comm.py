"""Communicator helpers (mirror of pylibwholegraph/torch/comm.py).

The unique id is minted by the group root and broadcast with torch.distributed, exactly like the
reference (comm.py:133-172); the id tensor is moved to the GPU only when the process group's
backend needs it (nccl), so the same code runs over gloo on a CPU box.
"""
import torch
import torch.distributed as dist

from .. import binding as wmb
from .utils import (str_to_wmb_wholememory_distributed_backend_type, str_to_wmb_wholememory_location,
                    str_to_wmb_wholememory_memory_type, wholememory_distributed_backend_type_to_str)

global_communicators = {}
local_node_communicator = None
local_device_communicator = None

all_comm_world_rank = 0
all_comm_world_size = 1
all_comm_local_rank = 0
all_comm_local_size = 1


def reset_communicators():
    global all_comm_world_rank, all_comm_world_size, all_comm_local_rank, all_comm_local_size
    global global_communicators, local_node_communicator, local_device_communicator, local_mnnvl_communicator
    global_communicators = {}
    local_node_communicator = None
    local_device_communicator = None
    local_mnnvl_communicator = None
    all_comm_world_rank = 0
    all_comm_world_size = 1
    all_comm_local_rank = 0
    all_comm_local_size = 1


def set_world_info(world_rank: int, world_size: int, local_rank: int, local_size: int):
    global all_comm_world_rank, all_comm_world_size, all_comm_local_rank, all_comm_local_size
    all_comm_world_rank = world_rank
    all_comm_world_size = world_size
    all_comm_local_rank = local_rank
    all_comm_local_size = local_size


class WholeMemoryCommunicator(object):
    """WholeMemory Communicator; create through create_group_communicator / get_global_communicator."""

    def __init__(self, wmb_comm: wmb.PyWholeMemoryComm):
        super().__init__()
        self.wmb_comm = wmb_comm

    def get_rank(self):
        return self.wmb_comm.get_rank()

    def get_size(self):
        return self.wmb_comm.get_size()

    def get_clique_info(self):
        return self.wmb_comm.get_clique_info()

    def barrier(self):
        return self.wmb_comm.barrier()

    def support_type_location(self, memory_type: str, memory_location: str):
        return self.wmb_comm.support_type_location(str_to_wmb_wholememory_memory_type(memory_type),
                                                   str_to_wmb_wholememory_location(memory_location))

    def destroy(self):
        wmb.destroy_communicator(self.wmb_comm)
        self.wmb_comm = None

    @property
    def distributed_backend(self):
        return wholememory_distributed_backend_type_to_str(self.wmb_comm.get_distributed_backend())

    @distributed_backend.setter
    def distributed_backend(self, value):
        self.wmb_comm.set_distributed_backend(str_to_wmb_wholememory_distributed_backend_type(value))


def _broadcast_uid(uid_th: torch.Tensor, root: int):
    if dist.get_backend() == "nccl":
        dev = uid_th.cuda()
        dist.broadcast(dev, root)
        uid_th.copy_(dev.cpu())
    else:
        dist.broadcast(uid_th, root)


def create_group_communicator(group_size: int = -1, comm_stride: int = 1):
    """24 ranks, group_size=4, comm_stride=2 -> [0,2,4,6], [1,3,5,7], [8,10,12,14], ... (reference comm.py:133)."""
    world_size = dist.get_world_size() if dist.is_initialized() else 1
    world_rank = dist.get_rank() if dist.is_initialized() else 0
    if group_size == -1:
        group_size = world_size
    strided_group_size = group_size * comm_stride
    assert world_size % strided_group_size == 0
    strided_group_count = world_size // strided_group_size
    strided_group_idx = world_rank // strided_group_size
    idx_in_strided_group = world_rank % strided_group_size
    inner_group_idx = idx_in_strided_group % comm_stride
    idx_in_group = idx_in_strided_group // comm_stride
    wm_uid = wmb.PyWholeMemoryUniqueID()
    for strided_group in range(strided_group_count):
        for inner_group in range(comm_stride):
            group_root_rank = strided_group * strided_group_size + inner_group
            tmp_wm_uid = wmb.create_unique_id() if world_rank == group_root_rank else wmb.PyWholeMemoryUniqueID()
            uid_th = tmp_wm_uid.as_tensor()
            if world_size > 1:
                _broadcast_uid(uid_th, group_root_rank)
            if strided_group_idx == strided_group and inner_group_idx == inner_group:
                wm_uid.as_tensor().copy_(uid_th)
    wm_comm = wmb.create_communicator(wm_uid, idx_in_group, group_size)
    return WholeMemoryCommunicator(wm_comm)


def split_communicator(comm: WholeMemoryCommunicator, color: int, key: int = 0):
    if not isinstance(color, int) or not isinstance(key, int):
        raise TypeError("color and key must be int")
    if color < 0:
        return None
    return WholeMemoryCommunicator(wmb.split_communicator(comm.wmb_comm, color, key))


def destroy_communicator(wm_comm: WholeMemoryCommunicator):
    if wm_comm is not None and wm_comm.wmb_comm is not None:
        wmb.destroy_communicator(wm_comm.wmb_comm)
        wm_comm.wmb_comm = None


def comm_set_distributed_backend(wm_comm: WholeMemoryCommunicator, distributed_backend: str):
    wmb.communicator_set_distributed_backend(wm_comm.wmb_comm,
                                             str_to_wmb_wholememory_distributed_backend_type(distributed_backend))


def get_global_communicator(distributed_backend="nccl"):
    global global_communicators, local_node_communicator, local_device_communicator
    if distributed_backend not in global_communicators:
        global_communicator = create_group_communicator()
        comm_set_distributed_backend(global_communicator, distributed_backend)
        global_communicators[distributed_backend] = global_communicator
        if distributed_backend == "nccl":
            if local_node_communicator is None and all_comm_local_size == all_comm_world_size:
                local_node_communicator = global_communicator
            if local_device_communicator is None and all_comm_world_size == 1:
                local_device_communicator = global_communicator
    return global_communicators[distributed_backend]


def get_local_node_communicator():
    global global_communicators, local_node_communicator, local_device_communicator
    if local_node_communicator is None:
        local_node_communicator = create_group_communicator(all_comm_local_size)
        if all_comm_local_size == all_comm_world_size:
            assert "nccl" not in global_communicators
            global_communicators["nccl"] = local_node_communicator
        if all_comm_local_size == 1:
            assert local_device_communicator is None
            local_device_communicator = local_node_communicator
    return local_node_communicator


def get_local_device_communicator():
    global global_communicators, local_node_communicator, local_device_communicator
    if local_device_communicator is None:
        local_device_communicator = create_group_communicator(1)
        if all_comm_local_size == 1:
            assert local_node_communicator is None
            local_node_communicator = local_device_communicator
        if all_comm_world_size == 1:
            assert "nccl" not in global_communicators
            global_communicators["nccl"] = local_device_communicator
    return local_device_communicator


local_mnnvl_communicator = None


def get_local_mnnvl_communicator():
    """Communicator over the ranks of this GPU's multi-node-NVLink clique (reference comm.py:257-279).  A single
    NVSwitch box reports no clique (wholememory_communicator_get_clique_info: is_in_clique = 0), so this raises the
    same RuntimeError as the reference does on non-MNNVL hardware."""
    global local_mnnvl_communicator
    if local_mnnvl_communicator is None:
        g_communicator = get_global_communicator()
        is_in_clique, _, _, _, clique_id, _ = g_communicator.get_clique_info()
        if not is_in_clique:
            raise RuntimeError("the gpu does not belong to any mnnvl domain,can not create local_mnnvl_communicator")
        local_mnnvl_communicator = split_communicator(g_communicator, clique_id)
    return local_mnnvl_communicator
