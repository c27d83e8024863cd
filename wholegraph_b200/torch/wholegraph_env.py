"""Allocator callbacks backed by torch (mirror of pylibwholegraph/torch/wholegraph_env.py:29-190).

Every temporary / output buffer an op needs is a torch tensor: temporaries ride torch's caching
allocator, outputs are returned to the caller as tensors.  The callbacks are ctypes closures over
the same four-function protocol as the reference (create_ctx / malloc / free / destroy_ctx).
"""
import ctypes
import glob
import importlib.util
import os
from typing import Union

import torch

from .. import _lib
from .. import binding as wmb
from .utils import torch_dtype_to_wholememory_dtype, wholememory_dtype_to_torch_dtype

default_wholegraph_env_context = None
_parked = []

# Native (C++) env functions: wholegraph_b200/csrc/torch_ext/torch_env.cpp, built in-tree into wholegraph_b200/lib/.
# With them an op call never re-enters the interpreter for its allocations (reference: the optional torch_cpp_ext,
# pylibwholegraph/torch/wholegraph_env.py:183-231).  Default when the module is built: a multi-hop sampling step on the
# full C5 graph takes 0.417 ms with them and 0.580 ms with the Python callbacks (B200, round 2; the reference's kernels
# behind the same Python callbacks: 0.517 ms).  WG_TORCH_NATIVE_ENV=0 keeps the ctypes closures, =1 insists on the module.
torch_cpp_ext_loaded = False
torch_cpp_ext_lib = None


def load_native_env(required: bool = False) -> bool:
    """Load the in-tree native env-function module; returns whether it is active."""
    global torch_cpp_ext_loaded, torch_cpp_ext_lib, default_wholegraph_env_context
    if torch_cpp_ext_loaded:
        return True
    if torch_cpp_ext_lib is not None:  # loaded earlier and switched off by unload_native_env(): switch it back on
        torch_cpp_ext_loaded = True
        _retire_default_context()
        return True
    lib_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib")
    found = sorted(glob.glob(os.path.join(lib_dir, "wholegraph_b200_torch_ext*.so")))
    if not found:
        if required:
            raise ImportError("wholegraph_b200_torch_ext is not built (run __graft_entry__.build())")
        return False
    spec = importlib.util.spec_from_file_location("wholegraph_b200_torch_ext", found[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch_cpp_ext_lib = mod
    torch_cpp_ext_loaded = True
    _retire_default_context()  # rebuilt on next use with the native table
    return True


def compile_cpp_extension():
    """Build (if needed) and activate the native env functions -- the reference's entry point of the same name
    (pylibwholegraph/torch/wholegraph_env.py:189-231) JIT-compiles its torch_cpp_ext; here the module is built
    in-tree by wholegraph_b200/csrc/torch_ext/build.sh."""
    import subprocess
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "torch_ext", "build.sh")
    subprocess.check_call(["bash", script])
    load_native_env(required=True)


def unload_native_env():
    """Back to the Python-callback env functions (the module stays imported)."""
    global torch_cpp_ext_loaded
    torch_cpp_ext_loaded = False
    _retire_default_context()


def _retire_default_context():
    """Stop handing out the current default table but keep it alive: callers may still hold its address."""
    global default_wholegraph_env_context
    if default_wholegraph_env_context is not None:
        _parked.append(default_wholegraph_env_context)
    default_wholegraph_env_context = None


def current_output_device() -> str:
    """Where op outputs are allocated: the current CUDA device (one place, so host-only tests can redirect it)."""
    return "cuda:%d" % torch.cuda.current_device()


def get_stream():
    cuda_stream = torch.cuda.current_stream()._as_parameter_
    return cuda_stream.value if cuda_stream.value is not None else 0


class TorchMemoryContext(object):
    """Owns one torch tensor allocated on behalf of the library."""
    _live = {}

    def __init__(self, native=None):
        self.tensor = None
        self.handle = 0
        if torch_cpp_ext_loaded if native is None else native:
            self._ext = torch_cpp_ext_lib
            self.handle = self._ext.create_output_context()
        else:
            self._ext = None
            TorchMemoryContext._live[id(self)] = self

    def get_c_context(self):
        return self.handle if self._ext is not None else id(self)

    def set_tensor(self, t):
        self.tensor = t

    def get_tensor(self):
        if self._ext is not None and self.handle != 0:
            self.tensor = self._ext.get_tensor_from_context(self.handle)
        return self.tensor

    def free_data(self):
        self.tensor = None
        if self._ext is not None and self.handle != 0:
            self._ext.free_context_data(self.handle)

    def free(self):
        self.tensor = None
        if self._ext is not None:
            if self.handle != 0:
                self._ext.destroy_output_context(self.handle)
                self.handle = 0
        else:
            TorchMemoryContext._live.pop(id(self), None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


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
    ctx = TorchMemoryContext(native=False)
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


class ExtContextWrapper(object):
    """The native module's static wholememory_env_func_t table."""

    def __init__(self, env_func: int):
        self.env_func = env_func

    def get_env_fns(self) -> int:
        return self.env_func


def create_current_env_context():
    if torch_cpp_ext_loaded:
        return ExtContextWrapper(torch_cpp_ext_lib.get_wholegraph_env_fns())
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


_native = os.environ.get("WG_TORCH_NATIVE_ENV", "")
if _native == "1":
    load_native_env(required=True)
elif _native != "0":
    try:
        load_native_env(required=False)
    except Exception as _e:  # a module built against another torch: the Python callbacks still work
        import warnings
        warnings.warn("wholegraph_b200: native env functions unavailable (%r); using the Python callbacks" % (_e,))
