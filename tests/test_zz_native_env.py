"""Native (C++) torch env functions -- wholegraph_b200/csrc/torch_ext/torch_env.cpp.

CPU part: the module loads from the tree, exposes the reference extension's six entry points
(pylibwholegraph/torch_cpp_ext/wholegraph_torch_ext.cpp:50-66), and its C callbacks follow the
create_ctx -> malloc(desc, kind, ctx) -> free(ctx) -> destroy_ctx protocol of
include/wholememory/env_func_ptrs.h for HOST allocations of every dtype.  GPU part: a sampling + gather +
append_unique pass with the native table must give the same results as the Python-callback table."""
import ctypes

import numpy as np
import pytest
import torch

from wholegraph_b200 import _lib
from wholegraph_b200.torch import wholegraph_env as wenv


@pytest.fixture()
def native():
    if not wenv.load_native_env():
        pytest.skip("wholegraph_b200_torch_ext is not built")
    yield wenv.torch_cpp_ext_lib


def _env_table(addr):
    return ctypes.cast(addr, ctypes.POINTER(_lib.EnvFns)).contents


def _desc(shape, dtype):
    d = _lib.TensorDescription()
    for i, s in enumerate(shape):
        d.sizes[i] = s
    stride = 1
    for i in reversed(range(len(shape))):
        d.strides[i] = stride
        stride *= shape[i]
    d.dim = len(shape)
    d.dtype = dtype
    d.storage_offset = 0
    return d


def test_module_surface(native):
    for name in ("get_wholegraph_env_fns", "get_stream", "create_output_context", "destroy_output_context",
                 "free_context_data", "get_tensor_from_context"):
        assert callable(getattr(native, name))
    assert native.get_wholegraph_env_fns() != 0
    assert native.get_wholegraph_env_fns() == wenv.get_wholegraph_env_fns()


DTYPES = [(1, torch.float32), (2, torch.float16), (3, torch.float64), (4, torch.bfloat16), (5, torch.int32), (6, torch.int64),
          (7, torch.int16), (8, torch.int8)]


@pytest.mark.parametrize("wm_dtype,th_dtype", DTYPES)
def test_temporary_protocol_host(native, wm_dtype, th_dtype):
    env = _env_table(native.get_wholegraph_env_fns())
    before = native.live_context_count()
    ctx = ctypes.c_void_p()
    env.temporary_fns.create_memory_context_fn(ctypes.byref(ctx), None)
    assert ctx.value and native.live_context_count() == before + 1
    d = _desc((5, 7), wm_dtype)
    ptr = env.temporary_fns.malloc_fn(ctypes.byref(d), 2, ctx, None)  # WHOLEMEMORY_MA_HOST
    assert ptr
    t = native.get_tensor_from_context(ctx.value)
    assert t.shape == (5, 7) and t.dtype == th_dtype and t.device.type == "cpu" and t.data_ptr() == ptr
    ctypes.memset(ptr, 0, t.numel() * t.element_size())
    assert int(t.to(torch.float64).abs().sum()) == 0
    env.temporary_fns.free_fn(ctx, None)
    assert native.get_tensor_from_context(ctx.value) is None
    env.temporary_fns.destroy_memory_context_fn(ctx, None)
    assert native.live_context_count() == before


def test_output_context_lifecycle(native):
    env = _env_table(native.get_wholegraph_env_fns())
    before = native.live_context_count()
    c = wenv.TorchMemoryContext()
    assert c.get_c_context() == c.handle != 0
    assert c.get_tensor() is None
    d = _desc((11,), 6)
    ptr = env.output_fns.malloc_fn(ctypes.byref(d), 2, c.get_c_context(), None)
    np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(11,))[:] = np.arange(11)
    t = c.get_tensor()
    assert t.tolist() == list(range(11))
    c.free()
    assert c.handle == 0 and native.live_context_count() == before
    assert t.tolist() == list(range(11))  # the tensor outlives its context
    # zero-size and bad requests
    c = wenv.TorchMemoryContext()
    d = _desc((0,), 5)
    env.output_fns.malloc_fn(ctypes.byref(d), 2, c.get_c_context(), None)
    assert c.get_tensor().numel() == 0
    d = _desc((4,), 0)  # WHOLEMEMORY_DT_UNKNOWN
    assert not env.output_fns.malloc_fn(ctypes.byref(d), 2, c.get_c_context(), None)
    d = _desc((4,), 1)
    assert not env.output_fns.malloc_fn(ctypes.byref(d), 0, c.get_c_context(), None)  # WHOLEMEMORY_MA_NONE
    del c
    assert native.live_context_count() == before


def test_python_callbacks_unaffected_after_unload():
    wenv.unload_native_env()
    c = wenv.TorchMemoryContext()
    assert c.get_c_context() == id(c)
    c.free()


@pytest.mark.gpu
def test_native_env_matches_python_callbacks_on_gpu():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "native_env_worker.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "native env worker OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_python_callback_table_protocol_host():
    """The DEFAULT env table (ctypes closures over torch) follows the same protocol; exercised here without a GPU
    through HOST allocations, exactly as the library calls it."""
    wenv.unload_native_env()
    env = _env_table(wenv.get_wholegraph_env_fns())
    ctx = ctypes.c_void_p()
    env.temporary_fns.create_memory_context_fn(ctypes.byref(ctx), None)
    assert ctx.value
    d = _desc((3, 4), 6)
    ptr = env.temporary_fns.malloc_fn(ctypes.byref(d), 2, ctx, None)
    assert ptr
    t = wenv.TorchMemoryContext._live[ctx.value].get_tensor()
    assert t.shape == (3, 4) and t.dtype == torch.int64 and t.data_ptr() == ptr
    env.temporary_fns.free_fn(ctx, None)
    assert wenv.TorchMemoryContext._live[ctx.value].get_tensor() is None
    env.temporary_fns.destroy_memory_context_fn(ctx, None)
    assert ctx.value not in wenv.TorchMemoryContext._live
    # caller-owned output context
    c = wenv.TorchMemoryContext()
    d = _desc((9,), 1)
    ptr = env.output_fns.malloc_fn(ctypes.byref(d), 2, c.get_c_context(), None)
    assert c.get_tensor().data_ptr() == ptr and c.get_tensor().dtype == torch.float32
    c.free()
    assert id(c) not in wenv.TorchMemoryContext._live
