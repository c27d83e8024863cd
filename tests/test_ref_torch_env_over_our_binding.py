"""The reference's Python allocation layer (pylibwholegraph/torch/wholegraph_env.py, loaded unchanged through compat/)
running on THIS repo's binding: its callbacks are registered through wmb.GlobalContextWrapper.create_context with the
reference's calling convention, and the resulting wholememory_env_func_t table is then driven the way the library drives it
(create_ctx -> malloc(desc, kind, ctx) -> free(ctx) -> destroy_ctx; output malloc into a caller-owned context).  This is the
flip side of tests/test_binding_surface.py (which runs this repo's callbacks through the reference's compiled module).
HOST allocations only, so it runs on CPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

from wholegraph_b200 import _lib

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/wholegraph_env.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_env():
    from compat_loader import load_reference_file
    return load_reference_file(REF, "_reference_wholegraph_env", package="pylibwholegraph.torch")


def _desc(shape, dtype):
    d = _lib.TensorDescription()
    stride = 1
    for i in reversed(range(len(shape))):
        d.sizes[i], d.strides[i] = shape[i], stride
        stride *= shape[i]
    d.dim, d.dtype, d.storage_offset = len(shape), dtype, 0
    return d


def test_reference_env_layer_allocates_through_this_binding(ref_env):
    addr = ref_env.get_wholegraph_env_fns()
    assert addr == ref_env.get_wholegraph_env_fns()                        # the default table is cached, as in the reference
    env = ctypes.cast(addr, ctypes.POINTER(_lib.EnvFns)).contents
    # temporary allocation: the library owns the context
    ctx = ctypes.c_void_p()
    env.temporary_fns.create_memory_context_fn(ctypes.byref(ctx), None)
    assert ctx.value
    for wm_dtype, th_dtype, shape in ((1, torch.float32, (5, 7)), (6, torch.int64, (11,)), (2, torch.float16, (3, 4)), (8, torch.int8, (0,))):
        d = _desc(shape, wm_dtype)
        ptr = env.temporary_fns.malloc_fn(ctypes.byref(d), 2, ctx, None)   # WHOLEMEMORY_MA_HOST
        n = int(np.prod(shape))
        assert ptr or n == 0
        if n:
            ctypes.memset(ptr, 0, n * torch.empty(0, dtype=th_dtype).element_size())
        env.temporary_fns.free_fn(ctx, None)
    env.temporary_fns.destroy_memory_context_fn(ctx, None)
    # output allocation: the caller owns a reference-layer TorchMemoryContext and passes id(obj), as wholegraph_ops.py does
    out = ref_env.TorchMemoryContext()
    d = _desc((9,), 5)                                                      # int32
    ptr = env.output_fns.malloc_fn(ctypes.byref(d), 2, out.get_c_context(), None)
    np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int32)), shape=(9,))[:] = np.arange(9)
    t = out.get_tensor()
    assert t.dtype == torch.int32 and t.tolist() == list(range(9)) and t.data_ptr() == ptr
    env.output_fns.free_fn(out.get_c_context(), None)
    assert out.get_tensor() is None
    out.free()


def test_reference_wrap_torch_tensor_produces_this_librarys_tensor(ref_env):
    import wholegraph_b200.binding as wmb
    t = torch.arange(60, dtype=torch.float32).reshape(6, 10)[:, 2:7]       # strides (10, 1), storage offset folded into the pointer
    w = ref_env.wrap_torch_tensor(t)
    assert isinstance(w, wmb.WrappedLocalTensor)
    d = _lib.lib.wholememory_tensor_get_tensor_description(ctypes.c_void_p(w.get_c_handle())).contents
    assert (d.dim, d.sizes[0], d.sizes[1], d.strides[0], d.strides[1], d.storage_offset, d.dtype) == (2, 6, 5, 10, 1, 0, int(wmb.DtFloat))
    assert _lib.lib.wholememory_tensor_get_data_pointer(ctypes.c_void_p(w.get_c_handle())) == t.data_ptr()
    none = ref_env.wrap_torch_tensor(None)
    assert _lib.lib.wholememory_tensor_get_tensor_description(ctypes.c_void_p(none.get_c_handle())).contents.dim == 0
