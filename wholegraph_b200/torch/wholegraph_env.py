"""Allocator callbacks backed by torch (mirror of pylibwholegraph/torch/wholegraph_env.py:29-190).

Every temporary / output buffer an op needs is a torch tensor: temporaries ride torch's caching
allocator, outputs are returned to the caller as tensors.  The callbacks are ctypes closures over
the same four-function protocol as the reference (create_ctx / malloc / free / destroy_ctx).
"""
import ctypes
from typing import Union

import torch

from .. import _lib
from .. import binding as wmb
from .utils import torch_dtype_to_wholememory_dtype, wholememory_dtype_to_torch_dtype

default_wholegraph_env_context = None


def get_stream():
    cuda_stream = torch.cuda.current_stream()._as_parameter_
    return cuda_stream.value if cuda_stream.value is not None else 0


class TorchMemoryContext(object):
    """Owns one torch tensor allocated on behalf of the library."""
    _live = {}

    def __init__(self):
        self.tensor = None
        TorchMemoryContext._live[id(self)] = self

    def get_c_context(self):
        return id(self)

    def set_tensor(self, t):
        self.tensor = t

    def get_tensor(self):
        return self.tensor

    def free_data(self):
        self.tensor = None

    def free(self):
        self.tensor = None
        TorchMemoryContext._live.pop(id(self), None)

    def __del__(self):
        TorchMemoryContext._live.pop(id(self), None)


def _ctx(handle):
    return TorchMemoryContext._live[handle]


def _torch_malloc(desc_ptr, malloc_type, memory_context, global_context):
    d = desc_ptr.contents
    shape = tuple(d.sizes[i] for i in range(d.dim))
    dtype = wholememory_dtype_to_torch_dtype(d.dtype)
    if malloc_type == wmb.WholeMemoryMemoryAllocType.MatDevice:
        t = torch.empty(shape, dtype=dtype, device="cuda")
    elif malloc_type == wmb.WholeMemoryMemoryAllocType.MatHost:
        t = torch.empty(shape, dtype=dtype, device="cpu")
    else:
        t = torch.empty(shape, dtype=dtype, device="cpu", pin_memory=True)
    _ctx(memory_context).set_tensor(t)
    return t.data_ptr()


def _torch_free(memory_context, global_context):
    ctx = TorchMemoryContext._live.get(memory_context)
    if ctx is not None:
        ctx.free_data()


_temp_contexts = {}


def _torch_create_ctx(out_ctx, global_context):
    ctx = TorchMemoryContext()
    _temp_contexts[id(ctx)] = ctx  # library-owned until destroy
    out_ctx[0] = id(ctx)


def _torch_destroy_ctx(memory_context, global_context):
    ctx = _temp_contexts.pop(memory_context, None)
    if ctx is not None:
        ctx.free()


class GlobalContextWrapper(object):
    """Holds the ctypes closures alive and exposes the wholememory_env_func_t* as an int
    (reference: wmb.GlobalContextWrapper, wholememory_binding.pyx:366-436)."""

    def __init__(self):
        self._create = _lib.CREATE_CTX_FN(_torch_create_ctx)
        self._destroy = _lib.DESTROY_CTX_FN(_torch_destroy_ctx)
        self._malloc = _lib.MALLOC_FN(_torch_malloc)
        self._free = _lib.FREE_FN(_torch_free)
        self.env = _lib.EnvFns()
        self.env.temporary_fns.create_memory_context_fn = self._create
        self.env.temporary_fns.destroy_memory_context_fn = self._destroy
        self.env.temporary_fns.malloc_fn = self._malloc
        self.env.temporary_fns.free_fn = self._free
        self.env.temporary_fns.global_context = None
        self.env.output_fns.malloc_fn = self._malloc
        self.env.output_fns.free_fn = self._free
        self.env.output_fns.global_context = None

    def get_env_fns(self) -> int:
        return ctypes.addressof(self.env)


def create_current_env_context():
    return GlobalContextWrapper()


def get_wholegraph_env_fns(use_default=True) -> int:
    global default_wholegraph_env_context
    if default_wholegraph_env_context is None or not use_default:
        ctx = create_current_env_context()
        if use_default:
            default_wholegraph_env_context = ctx
        else:
            # caller keeps nothing: park it so the closures outlive the call
            _parked.append(ctx)
        return ctx.get_env_fns()
    return default_wholegraph_env_context.get_env_fns()


_parked = []


def wrap_torch_tensor(t: Union[torch.Tensor, None]) -> wmb.WrappedLocalTensor:
    py_desc = wmb.PyWholeMemoryTensorDescription()
    wm_t = wmb.WrappedLocalTensor()
    if t is None:
        return wm_t.wrap_tensor(py_desc, 0)
    py_desc.set_dtype(torch_dtype_to_wholememory_dtype(t.dtype))
    py_desc.set_storage_offset(0)
    py_desc.set_shape(tuple(t.shape))
    py_desc.set_stride(tuple(t.stride()))
    w = wm_t.wrap_tensor(py_desc, t.data_ptr())
    w._keepalive = t
    return w
