"""WholeMemoryEmbedding / WholeMemoryOptimizer / WholeMemoryEmbeddingModule
(mirror of pylibwholegraph/torch/embedding.py: optimizer :33-70, lookup autograd fn :213-243,
embedding :246-377, create :380-470, module :537-555, optimizer factory :558-588).

Cache policies exist as objects for API compatibility, but creating an embedding WITH one raises
NotImplementedError (out of scope for an all-HBM B200 box)."""
from typing import List, Union

import torch

from .. import binding as wmb
from .comm import (WholeMemoryCommunicator, get_global_communicator, get_local_device_communicator,
                   get_local_node_communicator)
from .tensor import WholeMemoryTensor
from .utils import (get_file_size, str_to_wmb_wholememory_access_type, str_to_wmb_wholememory_location,
                    str_to_wmb_wholememory_memory_type, str_to_wmb_wholememory_optimizer_type,
                    torch_dtype_to_wholememory_dtype)
from .wholegraph_env import current_output_device, get_stream, get_wholegraph_env_fns, wrap_torch_tensor


class WholeMemoryOptimizer(object):
    """One sparse optimizer (SGD / LazyAdam / AdaGrad / RMSProp) driving any number of WholeMemoryEmbeddings.
    Build it with create_wholememory_optimizer; call step(lr) once per training iteration on every rank."""

    def __init__(self, global_comm: WholeMemoryCommunicator):
        super().__init__()
        self.global_comm = global_comm
        self.wmb_opt = wmb.WholeMemoryOptimizer()
        self.embeddings = []

    def add_embedding(self, wm_embedding):
        """Attach the optimizer (allocates its state tensors, collectively).  An embedding takes one optimizer, once."""
        assert isinstance(wm_embedding, WholeMemoryEmbedding)
        if wm_embedding.wmb_optimizer is not None:
            raise ValueError("optimizer can only be set once.")
        self.wmb_opt.add_embedding(wm_embedding.wmb_embedding)
        wm_embedding.wmb_optimizer = self.wmb_opt
        wm_embedding.dummy_input.requires_grad_(True)  # lets autograd reach EmbeddingLookupFn.backward
        self.embeddings.append(wm_embedding)

    def step(self, lr: float):
        """Push every embedding's pending sparse gradients to their owners and update the rows, then barrier so that
        no rank reads a row another rank is still updating."""
        for pending in (e for e in self.embeddings if e.need_apply):
            pending.apply_gradients(lr)
        self.global_comm.barrier()


class WholeMemoryCachePolicy(object):
    """Python handle of a wholememory_embedding_cache_policy_t (kept for API compatibility, see the module docstring)."""

    def __init__(self, wmb_cache_policy: wmb.WholeMemoryCachePolicy):
        super().__init__()
        self.wmb_cache_policy = wmb_cache_policy


def create_wholememory_cache_policy(cache_comm: WholeMemoryCommunicator, *, memory_type: str = "chunked",
                                    memory_location: str = "cuda", access_type: str = "readonly", ratio: float = 0.5):
    policy = wmb.WholeMemoryCachePolicy()
    policy.create_policy(cache_comm.wmb_comm, str_to_wmb_wholememory_memory_type(memory_type),
                         str_to_wmb_wholememory_location(memory_location), str_to_wmb_wholememory_access_type(access_type), ratio)
    return WholeMemoryCachePolicy(policy)


def destroy_wholememory_cache_policy(cache_policy: WholeMemoryCachePolicy):
    cache_policy.wmb_cache_policy.destroy_policy()
    cache_policy.wmb_cache_policy = None


_BUILTIN_CACHE_COMM = {"all_devices": get_global_communicator, "local_node": get_local_node_communicator,
                       "local_device": get_local_device_communicator}


def create_builtin_cache_policy(builtin_cache_type: str, embedding_memory_type: str, embedding_memory_location: str,
                                access_type: str, cache_ratio: float, *, cache_memory_type: str = "",
                                cache_memory_location: str = ""):
    """Named cache layouts of the reference (pylibwholegraph/torch/embedding.py:122-211): "none" -> None; "all_devices",
    "local_node", "local_device" pick the communicator the cache is sharded over and the cache's memory type
    (embedding's type / chunked / continuous unless given).  Same argument checks and defaults; note that creating an
    embedding WITH a cache policy is refused by this build (the cache hides host-memory latency, tables here live in HBM)."""
    if embedding_memory_type not in ("continuous", "chunked", "distributed", "hierarchy"):
        raise ValueError(f"embedding_memory_type={embedding_memory_type} is not valid")
    if embedding_memory_location not in ("cpu", "cuda"):
        raise ValueError(f"embedding_memory_location={embedding_memory_location} is not valid")
    if builtin_cache_type == "none":
        return None
    if cache_memory_location not in ("", "cpu", "cuda"):
        raise ValueError(f"cache_memory_location is {cache_memory_location}, should be empty or cpu, cuda")
    if builtin_cache_type not in _BUILTIN_CACHE_COMM:
        raise ValueError(f"builtin_cache_type={builtin_cache_type} not supported, "
                         f"should be none, local_device, local_node or all_devices")
    if builtin_cache_type == "all_devices":
        if embedding_memory_location == "cuda":
            print("[WARNING] a device cache in front of device memory costs memory and is slower than no cache")
        memory_type = cache_memory_type or embedding_memory_type
    elif builtin_cache_type == "local_node":
        memory_type = cache_memory_type or "chunked"
    else:
        memory_type = "continuous"
    return create_wholememory_cache_policy(_BUILTIN_CACHE_COMM[builtin_cache_type](), memory_type=memory_type,
                                           memory_location=cache_memory_location or "cuda", access_type=access_type,
                                           ratio=cache_ratio)


class EmbeddingLookupFn(torch.autograd.Function):
    """autograd bridge: forward = gather; backward hands (indices, dL/d rows) to the embedding, which keeps them until
    WholeMemoryOptimizer.step.  `dummy_input` is a 1-element parameter whose only job is to make this node differentiable."""

    @staticmethod
    def forward(ctx, indice: torch.Tensor, dummy_input: torch.Tensor, wm_embedding, is_training: bool = False,
                force_dtype: Union[torch.dtype, None] = None):
        rows = wm_embedding.gather(indice, is_training=is_training, force_dtype=force_dtype)
        if is_training and wm_embedding.need_grad():
            ctx.wm_embedding = wm_embedding
            ctx.save_for_backward(indice, rows, dummy_input)
        return rows

    @staticmethod
    def backward(ctx, grad_outputs: torch.Tensor):
        indice, _rows, dummy_input = ctx.saved_tensors
        ctx.wm_embedding.add_gradients(indice, grad_outputs)
        ctx.wm_embedding = None
        return None, torch.zeros_like(dummy_input), None, None, None


class WholeMemoryEmbedding(object):
    """An [N, D] embedding table in WholeMemory, trainable once an optimizer is attached."""

    def __init__(self, wmb_embedding: wmb.PyWholeMemoryEmbedding, wmb_cache_policy: Union[WholeMemoryCachePolicy, None]):
        super().__init__()
        self.wmb_embedding = wmb_embedding
        self.wmb_cache_policy = wmb_cache_policy
        self.adjust_cache = wmb_cache_policy is not None
        self.wmb_optimizer = None
        self.dummy_input = torch.nn.Parameter(torch.zeros(1), requires_grad=False)
        # lazily created views
        self.embedding_tensor = None
        self.optimizer_states = {}
        # sparse gradients recorded by backward passes since the last apply
        self.sparse_indices, self.sparse_grads = [], []
        self.need_apply = False

    @property
    def shape(self):
        return self.get_embedding_tensor().shape

    def dim(self):
        return self.get_embedding_tensor().dim()

    def set_adjust_cache(self, adjust_cache: bool):
        self.adjust_cache = bool(adjust_cache) and self.wmb_cache_policy is not None

    def need_grad(self):
        """True for any live embedding, exactly like the reference (embedding.py:276-277): whether a backward pass really
        reaches add_gradients is decided by dummy_input.requires_grad, which only WholeMemoryOptimizer.add_embedding sets."""
        return self.wmb_embedding is not None

    def gather(self, indice: torch.Tensor, *, is_training: bool = False, force_dtype: Union[torch.dtype, None] = None):
        """rows[i, :] = table[indice[i], :] on the current CUDA device; with is_training and an optimizer attached the
        result requires grad and the embedding is marked as having gradients to apply at the next optimizer step."""
        assert indice.dim() == 1
        table = self.get_embedding_tensor()
        track = is_training and self.need_grad()
        rows = torch.empty([indice.shape[0], table.shape[1]], device=current_output_device(),
                           dtype=table.dtype if force_dtype is None else force_dtype, requires_grad=track)
        if track:
            self.need_apply = True
        wmb.EmbeddingGatherForward(self.wmb_embedding, wrap_torch_tensor(indice), wrap_torch_tensor(rows), self.adjust_cache,
                                   get_wholegraph_env_fns(), get_stream())
        return rows

    def add_gradients(self, indice: torch.Tensor, grad_outputs: torch.Tensor):
        self.sparse_indices.append(indice)
        self.sparse_grads.append(grad_outputs)

    def apply_gradients(self, lr: float):
        """One wholememory_embedding_gather_gradient_apply over everything recorded since the last call (collective)."""
        indices, grads = torch.cat(self.sparse_indices), torch.cat(self.sparse_grads)
        wmb.EmbeddingGatherGradientApply(self.wmb_embedding, wrap_torch_tensor(indices), wrap_torch_tensor(grads), self.adjust_cache,
                                         lr, get_wholegraph_env_fns(), get_stream())
        self.sparse_indices, self.sparse_grads = [], []
        self.need_apply = False

    def writeback_all_cache(self):
        self.wmb_embedding.writeback_all_cache(get_stream())

    def drop_all_cache(self):
        self.wmb_embedding.drop_all_cache(get_stream())

    def get_embedding_tensor(self):
        if self.embedding_tensor is None:
            self.embedding_tensor = WholeMemoryTensor(self.wmb_embedding.get_embedding_tensor())
        return self.embedding_tensor

    def get_optimizer_state_names(self):
        return self.wmb_embedding.get_optimizer_state_names()

    def get_optimizer_state(self, state_name):
        state = self.optimizer_states.get(state_name)
        if state is None:
            state = self.optimizer_states[state_name] = WholeMemoryTensor(self.wmb_embedding.get_optimizer_state(state_name))
        return state

    def save(self, file_prefix: str):
        """Checkpoint the embedding rows and every optimizer state (one part file per rank and tensor)."""
        self.get_embedding_tensor().to_file_prefix(file_prefix + "_embedding_tensor")
        for state_name in self.get_optimizer_state_names():
            self.get_optimizer_state(state_name).to_file_prefix(file_prefix + "_" + state_name)

    def load(self, file_prefix: str, *, ignore_embedding: bool = False, part_count: Union[int, None] = None):
        if ignore_embedding is False:
            self.get_embedding_tensor().from_file_prefix(file_prefix + "_embedding_tensor", part_count)
        for state_name in self.get_optimizer_state_names():
            self.get_optimizer_state(state_name).from_file_prefix(file_prefix + "_" + state_name, part_count)


def create_embedding(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str, dtype: torch.dtype,
                     sizes: List[int], *, cache_policy: Union[WholeMemoryCachePolicy, None] = None,
                     embedding_entry_partition: Union[List[int], None] = None, random_init: bool = False,
                     gather_sms: int = -1, round_robin_size: int = 0):
    """Create a [sizes[0], sizes[1]] embedding table row-sharded over comm (collective; ends with a barrier).

    embedding_entry_partition[i] = rows owned by rank i; it is dropped (with a note) when a cache policy is given, and it
    switches round-robin sharding off.  random_init fills this rank's rows with Xavier-uniform values.  gather_sms limits
    the SMs a gather may use (-1 = all)."""
    assert len(sizes) == 2
    if embedding_entry_partition is not None:
        if cache_policy is not None:
            print("[wholegraph_b200] a cache policy decides the row partition itself: embedding_entry_partition dropped")
            embedding_entry_partition = None
        elif round_robin_size != 0:
            print("[wholegraph_b200] an explicit embedding_entry_partition excludes round-robin sharding: round_robin_size set to 0")
            round_robin_size = 0
    desc = wmb.PyWholeMemoryTensorDescription()
    desc.set_dtype(torch_dtype_to_wholememory_dtype(dtype))
    desc.set_shape(sizes)
    desc.set_stride([sizes[1], 1])
    policy = cache_policy.wmb_cache_policy if cache_policy is not None else wmb.create_non_cache_policy()
    handle = wmb.create_embedding(desc, comm.wmb_comm, str_to_wmb_wholememory_memory_type(memory_type),
                                  str_to_wmb_wholememory_location(memory_location), policy,
                                  embedding_entry_partition=embedding_entry_partition, user_defined_sms=gather_sms,
                                  round_robin_size=round_robin_size)
    wm_embedding = WholeMemoryEmbedding(handle, cache_policy)
    if random_init is True:
        my_rows, _first_row = wm_embedding.get_embedding_tensor().get_local_tensor()
        torch.nn.init.xavier_uniform_(my_rows)
    comm.barrier()
    return wm_embedding


def create_embedding_from_filelist(comm: WholeMemoryCommunicator, memory_type: str, memory_location: str,
                                   filelist: Union[List[str], str], dtype: torch.dtype, last_dim_size: int, *,
                                   cache_policy: Union[WholeMemoryCachePolicy, None] = None,
                                   embedding_entry_partition: Union[List[int], None] = None, gather_sms: int = -1,
                                   round_robin_size: int = 0):
    """Create an embedding sized from, and filled with, raw row-major binary files of [rows, last_dim_size] `dtype`
    (reference: pylibwholegraph/torch/embedding.py:462-524)."""
    if isinstance(filelist, str):
        filelist = [filelist]
    assert last_dim_size > 0
    # the same two rules as create_embedding, applied here as well because the file load below takes round_robin_size too
    if embedding_entry_partition is not None and cache_policy is not None:
        embedding_entry_partition = None
    if embedding_entry_partition is not None:
        round_robin_size = 0
    row_bytes = torch.tensor([], dtype=dtype).element_size() * last_dim_size
    total_bytes = 0
    for filename in filelist:
        file_size = get_file_size(filename)
        if file_size % row_bytes != 0:
            raise ValueError("File %s size is %d not mutlple of %d" % (filename, file_size, row_bytes))
        total_bytes += file_size
    wm_embedding = create_embedding(comm, memory_type, memory_location, dtype, [total_bytes // row_bytes, last_dim_size],
                                    cache_policy=cache_policy, embedding_entry_partition=embedding_entry_partition,
                                    gather_sms=gather_sms, round_robin_size=round_robin_size)
    wm_embedding.get_embedding_tensor().from_filelist(filelist, round_robin_size)
    return wm_embedding


def destroy_embedding(wm_embedding: WholeMemoryEmbedding):
    wm_embedding.wmb_embedding.destroy_embedding()
    wm_embedding.wmb_embedding = None


class WholeMemoryEmbeddingModule(torch.nn.Module):
    """nn.Module face of a WholeMemoryEmbedding: forward(indices) -> rows, differentiable while the module is in
    training mode and the embedding has an optimizer."""

    def __init__(self, wm_embedding: WholeMemoryEmbedding):
        super().__init__()
        self.wm_embedding = wm_embedding
        self.embedding_gather_fn = EmbeddingLookupFn.apply

    def forward(self, indice: torch.Tensor, force_dtype: Union[torch.dtype, None] = None):
        emb = self.wm_embedding
        return self.embedding_gather_fn(indice, emb.dummy_input, emb, self.training, force_dtype)


def create_wholememory_optimizer(embeddings: Union[WholeMemoryEmbedding, List[WholeMemoryEmbedding]], optimizer_type: str,
                                 param_dict: dict, global_comm: Union[WholeMemoryCommunicator, None] = None):
    """optimizer_type: "sgd" | "adam" (LazyAdam; adam_w=1 for AdamW) | "adagrad" | "rmsprop"; param_dict holds the float
    parameters by the reference's names (weight_decay, epsilon, beta1, beta2, adam_w, alpha).  global_comm (optional here,
    the global communicator by default) is what WholeMemoryOptimizer.step barriers on."""
    wm_optimizer = WholeMemoryOptimizer(get_global_communicator() if global_comm is None else global_comm)
    wm_optimizer.wmb_opt.create_optimizer(str_to_wmb_wholememory_optimizer_type(optimizer_type), param_dict)
    for emb in ([embeddings] if isinstance(embeddings, WholeMemoryEmbedding) else embeddings):
        wm_optimizer.add_embedding(emb)
    return wm_optimizer


def destroy_wholememory_optimizer(optimizer: WholeMemoryOptimizer):
    optimizer.wmb_opt.destroy_optimizer()
    optimizer.wmb_opt = None
