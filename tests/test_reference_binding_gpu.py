"""Drop-in proof: the REFERENCE's own cython binding (wholememory_binding.pyx, unmodified, cythonized against THIS repo's
headers and linked to THIS repo's libwholegraph.so by oracle/build_ref_binding.sh) drives the sm_100a kernels.
The env-function callbacks below are a transcription of what pylibwholegraph/torch/wholegraph_env.py registers."""
import ctypes
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIND_DIR = os.path.join(ROOT, "oracle", "_ref", "refbinding")
HAVE = bool(glob.glob(os.path.join(BIND_DIR, "wholememory_binding*.so")))


_IMMORTAL = []


class _Ctx(object):
    def __init__(self):
        self.tensor = None


@pytest.mark.skipif(not HAVE, reason="oracle/_ref/refbinding not built (needs /root/reference + cython at build time)")
def test_reference_cython_binding_on_our_library():
    import torch
    sys.path.insert(0, BIND_DIR)
    import wholememory_binding as rwmb  # the reference's module
    from oracle import oracle as O

    torch.cuda.set_device(0)
    rwmb.init(0)
    comm = rwmb.create_communicator(rwmb.create_unique_id(), 0, 1)
    assert comm.get_rank() == 0 and comm.get_size() == 1

    # --- env functions exactly as the reference's torch layer provides them (Python callbacks re-entered from C)
    def create_ctx(global_context):
        return _Ctx()

    def destroy_ctx(memory_context, global_context):
        memory_context.tensor = None

    def malloc_fn(tensor_desc, malloc_type, memory_context, global_context):
        dt = {rwmb.WholeMemoryDataType.DtFloat: torch.float32, rwmb.WholeMemoryDataType.DtInt: torch.int32,
              rwmb.WholeMemoryDataType.DtInt64: torch.int64, rwmb.WholeMemoryDataType.DtInt8: torch.int8,
              rwmb.WholeMemoryDataType.DtHalf: torch.float16}[tensor_desc.dtype]
        if malloc_type.get_type() == rwmb.WholeMemoryMemoryAllocType.MatDevice:
            t = torch.empty(tensor_desc.shape, dtype=dt, device="cuda")
        else:
            t = torch.empty(tensor_desc.shape, dtype=dt, pin_memory=malloc_type.get_type() == rwmb.WholeMemoryMemoryAllocType.MatPinned)
        memory_context.tensor = t
        return t.data_ptr()

    def free_fn(memory_context, global_context):
        memory_context.tensor = None

    gctx = object()
    env = rwmb.GlobalContextWrapper()
    env.create_context(create_ctx, destroy_ctx, malloc_fn, free_fn, gctx, malloc_fn, free_fn, gctx)
    # The reference keeps its GlobalContextWrapper in a module global for the life of the process
    # (torch/wholegraph_env.py:29-40); its __dealloc__ (wholememory_binding.pyx:387-397) raises inside tp_dealloc
    # ("no attribute 'self'"), and pytest's unraisable-exception hook then holds a dangling object -> SIGSEGV at
    # session end.  Same lifetime here: never collected.
    _IMMORTAL.append(env)
    ctypes.pythonapi.Py_IncRef(ctypes.py_object(env))

    def wrap(t):
        d = rwmb.PyWholeMemoryTensorDescription()
        d.set_dtype({torch.float32: rwmb.WholeMemoryDataType.DtFloat, torch.int64: rwmb.WholeMemoryDataType.DtInt64,
                     torch.float16: rwmb.WholeMemoryDataType.DtHalf, torch.int32: rwmb.WholeMemoryDataType.DtInt}[t.dtype])
        d.set_storage_offset(0)
        d.set_shape(tuple(t.shape))
        d.set_stride(tuple(t.stride()))
        return rwmb.WrappedLocalTensor().wrap_tensor(d, t.data_ptr())

    def from_dlpack(dp):
        return torch.utils.dlpack.from_dlpack(dp.__dlpack__())

    rows, cols = 5000, 96
    rng = np.random.default_rng(5)
    for mt in (rwmb.WholeMemoryMemoryType.MtContinuous, rwmb.WholeMemoryMemoryType.MtChunked, rwmb.WholeMemoryMemoryType.MtDistributed):
        wm = rwmb.create_wholememory_matrix(rwmb.WholeMemoryDataType.DtFloat, rows, cols, -1, comm, mt,
                                            rwmb.WholeMemoryMemoryLocation.MlDevice)
        assert wm.shape == (rows, cols) and wm.get_local_entry_count() == rows and wm.get_local_entry_start() == 0
        local, off = wm.get_local_tensor(from_dlpack, rwmb.WholeMemoryMemoryLocation.MlDevice, 0)  # DLPack export path
        assert off == 0 and tuple(local.shape) == (rows, cols)
        host = rng.standard_normal((rows, cols)).astype(np.float32)
        # scatter through the reference binding, then read the table both via the mapped view and via gather
        src = torch.from_numpy(host).cuda()
        all_idx = torch.arange(rows, dtype=torch.int64, device="cuda")
        rwmb.wholememory_scatter_op(wrap(src), wrap(all_idx), wm, env.get_env_fns(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(local, src)
        idx = rng.integers(0, rows, size=3333).astype(np.int64)
        idx[3] = -1
        out = torch.full((3333, cols), 9.0, dtype=torch.float16, device="cuda")  # fp32 table -> fp16 output
        idx_t = torch.from_numpy(idx).cuda()  # WrappedLocalTensor keeps a pointer only: hold the tensor across the call
        rwmb.wholememory_gather_op(wm, wrap(idx_t), wrap(out), env.get_env_fns(),
                                   torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        exp = np.full((3333, cols), 9.0, dtype=np.float16)
        O.gather(host, O.DT_FLOAT, idx, O.DT_HALF, out=exp)
        assert out.cpu().numpy().tobytes() == exp.tobytes()
        rwmb.destroy_wholememory_tensor(wm)

    # allocator plumbing self-test op through the reference's Python-callback env functions
    inp = torch.arange(16, dtype=torch.float32, device="cuda")
    fixed = torch.zeros(5, 16, device="cuda")
    c_dev, c_pin, c_host = _Ctx(), _Ctx(), _Ctx()
    rwmb.wholememory_env_test_cython_op(wrap(inp), wrap(fixed), id(c_dev), id(c_pin), id(c_host), 5, env.get_env_fns(),
                                        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    expect = inp.unsqueeze(0) + torch.arange(5, dtype=torch.float32, device="cuda").unsqueeze(1)
    assert torch.equal(fixed, expect)
    assert torch.equal(c_dev.tensor, expect) and torch.equal(c_pin.tensor.cuda(), expect) and torch.equal(c_host.tensor.cuda(), expect)

    assert rwmb.py_get_wholememory_tensor_count() >= 0
    rwmb.destroy_communicator(comm)
    # No rwmb.finalize() here: the reference module is linked to the SAME libwholegraph.so the ctypes binding has open, and
    # wholememory_finalize destroys every communicator of the process (reference initialize.cpp:73-77) -- including the one
    # tests/gpu_utils.py shares across test files.  (Round 1's GPU suite died of exactly that; the library now also refuses
    # stale handles, tests/test_stale_handles.py.)  finalize() through the reference binding is exercised in a process of
    # its own by tests/ref_binding_worker.py.
