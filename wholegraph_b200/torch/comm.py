"""Communicator helpers with the public names of pylibwholegraph/torch/comm.py.

How a communicator comes to exist here: the root rank of each group mints a 128-byte unique id, the id travels to the
other members over torch.distributed (staged on the GPU only when the process group's backend is nccl, so the same code
runs over gloo on a CPU box), and every member then joins the library's own AF_UNIX bootstrap under that id.
The well-known communicators (global per distributed backend, per node, per device, per MNNVL clique) are created on first
use and cached in one registry; when two of them cover the same ranks they are the same object, as in the reference
(comm.py:197-255).
"""
import torch.distributed as dist

from .. import binding as wmb
from .utils import (str_to_wmb_wholememory_distributed_backend_type, str_to_wmb_wholememory_location,
                    str_to_wmb_wholememory_memory_type, wholememory_distributed_backend_type_to_str)


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
        destroy_communicator(self)

    @property
    def distributed_backend(self):
        return wholememory_distributed_backend_type_to_str(self.wmb_comm.get_distributed_backend())

    @distributed_backend.setter
    def distributed_backend(self, value):
        self.wmb_comm.set_distributed_backend(str_to_wmb_wholememory_distributed_backend_type(value))


class _Registry(object):
    """World layout as told by init(), plus the cached well-known communicators."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.world_rank, self.world_size, self.local_rank, self.local_size = 0, 1, 0, 1
        self.by_backend = {}  # distributed backend name -> communicator over all ranks
        self.node = None      # ranks of this node
        self.device = None    # this rank alone
        self.mnnvl = None     # ranks of this GPU's multi-node-NVLink clique

    def adopt(self, comm, ranks_covered: int):
        """A freshly made communicator over `ranks_covered` ranks also serves every well-known role of that extent."""
        if ranks_covered == self.world_size and "nccl" not in self.by_backend:
            self.by_backend["nccl"] = comm
        if ranks_covered == self.local_size and self.node is None:
            self.node = comm
        if ranks_covered == 1 and self.device is None:
            self.device = comm


_reg = _Registry()


def reset_communicators():
    _reg.clear()


def set_world_info(world_rank: int, world_size: int, local_rank: int, local_size: int):
    _reg.world_rank, _reg.world_size = world_rank, world_size
    _reg.local_rank, _reg.local_size = local_rank, local_size


def _share_unique_id(uid, root: int):
    """Broadcast the bytes of `uid` from `root` over torch.distributed, in place."""
    buf = uid.as_tensor()
    if dist.get_backend() == "nccl":
        staged = buf.cuda()
        dist.broadcast(staged, root)
        buf.copy_(staged.cpu())
    else:
        dist.broadcast(buf, root)


def create_group_communicator(group_size: int = -1, comm_stride: int = 1):
    """Partition the world into communicators of `group_size` ranks whose members are `comm_stride` apart.

    24 ranks, group_size=4, comm_stride=2 -> [0,2,4,6], [1,3,5,7], [8,10,12,14], [9,11,13,15], ...  (reference comm.py:133-168).
    Collective over the whole process group: every rank takes part in every group's id broadcast."""
    have_pg = dist.is_initialized()
    world_size = dist.get_world_size() if have_pg else 1
    me = dist.get_rank() if have_pg else 0
    if group_size == -1:
        group_size = world_size
    block = group_size * comm_stride          # consecutive ranks that hold `comm_stride` interleaved groups
    assert world_size % block == 0
    my_block, in_block = divmod(me, block)
    my_lane, my_index = in_block % comm_stride, in_block // comm_stride
    mine = wmb.PyWholeMemoryUniqueID()
    for root in (b * block + lane for b in range(world_size // block) for lane in range(comm_stride)):
        uid = wmb.create_unique_id() if me == root else wmb.PyWholeMemoryUniqueID()
        if world_size > 1:
            _share_unique_id(uid, root)
        if root == my_block * block + my_lane:
            mine.as_tensor().copy_(uid.as_tensor())
    return WholeMemoryCommunicator(wmb.create_communicator(mine, my_index, group_size))


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
    """All ranks; one communicator per distributed backend name."""
    comm = _reg.by_backend.get(distributed_backend)
    if comm is None:
        comm = create_group_communicator()
        comm_set_distributed_backend(comm, distributed_backend)
        _reg.by_backend[distributed_backend] = comm
        if distributed_backend == "nccl":
            _reg.adopt(comm, _reg.world_size)
    return comm


def get_local_node_communicator():
    """The ranks of this node."""
    if _reg.node is None:
        _reg.adopt(create_group_communicator(_reg.local_size), _reg.local_size)
    return _reg.node


def get_local_device_communicator():
    """This rank alone."""
    if _reg.device is None:
        _reg.adopt(create_group_communicator(1), 1)
    return _reg.device


def get_local_mnnvl_communicator():
    """The ranks of this GPU's multi-node-NVLink clique (reference comm.py:257-279).  A single NVSwitch box reports no
    clique (wholememory_communicator_get_clique_info: is_in_clique = 0), so this raises the same RuntimeError as the
    reference does on non-MNNVL hardware."""
    if _reg.mnnvl is None:
        everyone = get_global_communicator()
        is_in_clique, _, _, _, clique_id, _ = everyone.get_clique_info()
        if not is_in_clique:
            raise RuntimeError("the gpu does not belong to any mnnvl domain,can not create local_mnnvl_communicator")
        _reg.mnnvl = split_communicator(everyone, clique_id)
    return _reg.mnnvl
