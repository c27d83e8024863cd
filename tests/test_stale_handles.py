"""Handles that outlived their object are refused at the C boundary (csrc/registry.cpp) instead of being dereferenced.

Round 1's GPU suite died with SIGSEGV because one test called wholememory_finalize() -- which destroys every communicator
of the process, like the reference's (cpp/src/wholememory/initialize.cpp:73-77) -- and a later test file passed the
communicator it had cached to wholememory_create_embedding.  The crash reproduces without a GPU; these tests replay it
(first with this repo's binding alone, then through the reference's own cython module the way the GPU suite hit it) and
walk every family of entry point with dead handles.  They run in child processes: a regression here is a segfault."""
import glob
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIND_DIR = os.path.join(ROOT, "oracle", "_ref", "refbinding")
HAVE_REF_BINDING = bool(glob.glob(os.path.join(BIND_DIR, "wholememory_binding*.so")))


def _run(body):
    src = "import sys\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body)
    p = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, "child rc=%d\nstdout:\n%s\nstderr:\n%s" % (p.returncode, p.stdout[-3000:], p.stderr[-3000:])
    assert "ALL-OK" in p.stdout, p.stdout[-3000:]


PRELUDE = """
    import ctypes
    import wholegraph_b200.binding as wmb
    from wholegraph_b200._lib import lib
    INVALID_INPUT = 6  # WHOLEMEMORY_INVALID_INPUT (include/wholememory/wholememory.h)

    def refused(fn, *a):
        try:
            fn(*a)
        except ValueError as e:  # the binding maps INVALID_INPUT to ValueError like the reference's check_wholememory_error_code
            assert "nvalid input" in str(e), e
            return True
        raise AssertionError("%s accepted a stale handle" % getattr(fn, "__name__", fn))

    def desc(rows, cols):
        d = wmb.PyWholeMemoryTensorDescription()
        d.set_dtype(wmb.DtFloat); d.set_shape((rows, cols)); d.set_stride((cols, 1))
        return d

    wmb.init(0, wmb.WholeMemoryLogLevel.LevFatal)
"""


def test_communicators_and_raw_tensors_without_a_gpu():
    """CPU: WholeMemory cannot be allocated without a device (no CPU fallback), so this covers communicators,
    raw-pointer tensors, views that outlive their root, and optimizers."""
    _run(PRELUDE + """
    import torch
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    assert comm.get_rank() == 0
    wmb.finalize()
    refused(comm.get_rank); refused(comm.get_size); refused(comm.barrier)
    assert comm.support_type_location(wmb.MtChunked, wmb.MlHost) is False   # bool API: a dead communicator supports nothing
    refused(wmb.destroy_communicator, comm)
    refused(wmb.malloc, 1 << 20, comm, wmb.MtDistributed, wmb.MlDevice, 64)
    refused(wmb.create_embedding, desc(64, 8), comm, wmb.MtChunked, wmb.MlHost, wmb.create_non_cache_policy())   # the round-1 crash
    assert lib.wholememory_communicator_barrier(ctypes.c_void_p(0xdeadbeef0)) == INVALID_INPUT       # never issued
    assert lib.wholememory_destroy_tensor(ctypes.c_void_p(0xdeadbeef0)) == INVALID_INPUT
    # raw-pointer tensors do not depend on a communicator: still alive after finalize, dead after destroy
    buf = torch.zeros(16, 8)
    d = desc(16, 8)
    root = ctypes.c_void_p()
    assert lib.wholememory_make_tensor_from_pointer(ctypes.byref(root), ctypes.c_void_p(buf.data_ptr()), ctypes.byref(d.tensor_description)) == 0
    starts = (ctypes.c_int64 * 2)(0, 0); ends = (ctypes.c_int64 * 2)(8, 8)
    view = ctypes.c_void_p()
    assert lib.wholememory_tensor_get_subtensor(root, starts, ends, ctypes.byref(view)) == 0
    lib.wholememory_tensor_get_data_pointer.restype = ctypes.c_void_p
    assert lib.wholememory_tensor_get_data_pointer(view) == buf.data_ptr()
    assert lib.wholememory_destroy_tensor(root) == 0
    n = ctypes.c_size_t()
    assert lib.wholememory_tensor_get_local_entry_count(ctypes.byref(n), view) == INVALID_INPUT      # a view that outlived its root
    assert lib.wholememory_tensor_get_data_pointer(view) is None
    assert lib.wholememory_tensor_get_data_pointer(root) is None
    assert lib.wholememory_destroy_tensor(view) == 0 and lib.wholememory_destroy_tensor(view) == INVALID_INPUT
    o = ctypes.c_void_p()
    assert lib.wholememory_create_embedding_optimizer(ctypes.byref(o), 1) == 0
    lib.wholememory_destroy_embedding_optimizer.restype = None
    lib.wholememory_destroy_embedding_optimizer(o)
    lib.wholememory_destroy_embedding_optimizer(o)                          # twice: ignored, not a double free
    v = ctypes.c_float(0.5)
    assert lib.wholememory_optimizer_set_parameter(o, b"weight_decay", ctypes.byref(v)) == INVALID_INPUT
    # the library is usable again afterwards
    wmb.init(0, wmb.WholeMemoryLogLevel.LevFatal)
    comm2 = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    assert comm2.get_size() == 1
    wmb.destroy_communicator(comm2)
    print("ALL-OK")
    """)


@pytest.mark.gpu
def test_every_entry_family_refuses_handles_after_finalize():
    _run(PRELUDE + """
    comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    t = wmb.create_wholememory_matrix(wmb.DtFloat, 64, 8, -1, comm, wmb.MtChunked, wmb.MlHost)
    h = t.get_wholememory_handle()
    sub = t.get_sub_tensor((0, 0), (32, 8))
    opt = wmb.create_optimizer(wmb.OptSgd, {})
    assert comm.get_rank() == 0 and h.get_local_memory()[1] == 64 * 8 * 4
    c_comm, c_h, c_t = comm.comm_id, h.wholememory_handle, t.wholememory_tensor

    wmb.finalize()   # destroys the communicator, which takes the handle's memory with it

    # communicator family
    refused(comm.get_rank); refused(comm.get_size); refused(comm.barrier)
    assert comm.support_type_location(wmb.MtChunked, wmb.MlHost) is False   # bool API: a dead communicator supports nothing
    refused(wmb.destroy_communicator, comm)
    refused(wmb.create_wholememory_matrix, wmb.DtFloat, 64, 8, -1, comm, wmb.MtChunked, wmb.MlHost)
    refused(wmb.create_embedding, desc(64, 8), comm, wmb.MtChunked, wmb.MlHost, wmb.create_non_cache_policy())   # the round-1 crash
    # handle family
    refused(h.get_local_memory); refused(h.get_global_pointer); refused(h.get_global_reference); refused(h.get_rank_memory, 0)
    for f in (lib.wholememory_get_total_size, lib.wholememory_get_data_granularity):
        f.restype = ctypes.c_size_t
        assert f(c_h) == 0
    assert lib.wholememory_get_memory_type(c_h) == 0 and lib.wholememory_get_memory_location(c_h) == 0
    assert lib.wholememory_free(c_h) == INVALID_INPUT
    assert lib.wholememory_load_from_file(c_h, 0, 32, 32, None, 0, 0) == INVALID_INPUT
    assert lib.wholememory_store_to_file(c_h, 0, 32, 32, b"/tmp/never") == INVALID_INPUT
    # tensor family: the tensor OBJECT is alive, its handle is not
    lib.wholememory_tensor_get_memory_handle.restype = ctypes.c_void_p
    lib.wholememory_tensor_get_data_pointer.restype = ctypes.c_void_p
    lib.wholememory_tensor_get_root.restype = ctypes.c_void_p
    assert lib.wholememory_tensor_get_memory_handle(c_t) is None
    assert lib.wholememory_tensor_get_data_pointer(c_t) is None
    assert lib.wholememory_tensor_get_root(c_t) is None
    n = ctypes.c_size_t()
    assert lib.wholememory_tensor_get_local_entry_count(ctypes.byref(n), c_t) == INVALID_INPUT
    assert lib.wholememory_tensor_get_local_entry_start(ctypes.byref(n), c_t) == INVALID_INPUT
    out = ctypes.c_void_p()
    assert lib.wholememory_tensor_map_local_tensor(c_t, ctypes.byref(out)) == INVALID_INPUT
    # ops on a dead table (the operands are live raw-pointer tensors)
    import numpy as np, torch
    from wholegraph_b200.torch.wholegraph_env import get_wholegraph_env_fns, wrap_torch_tensor
    idx = torch.zeros(4, dtype=torch.int64); dense = torch.zeros(4, 8)
    refused(wmb.wholememory_gather_op, t, wrap_torch_tensor(idx), wrap_torch_tensor(dense), get_wholegraph_env_fns(), 0)
    refused(wmb.wholememory_scatter_op, wrap_torch_tensor(dense), wrap_torch_tensor(idx), t, get_wholegraph_env_fns(), 0)
    # destroying what is left succeeds without touching freed memory, and only once
    wmb.destroy_wholememory_tensor(sub)
    wmb.destroy_wholememory_tensor(t)
    assert lib.wholememory_destroy_tensor(c_t) == INVALID_INPUT
    assert lib.wholememory_destroy_tensor(ctypes.c_void_p(0xdeadbeef0)) == INVALID_INPUT   # never issued by the library
    assert lib.wholememory_communicator_barrier(ctypes.c_void_p(0xdeadbeef0)) == INVALID_INPUT
    # the library is usable again afterwards
    wmb.init(0, wmb.WholeMemoryLogLevel.LevFatal)
    comm2 = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    t2 = wmb.create_wholememory_matrix(wmb.DtFloat, 64, 8, -1, comm2, wmb.MtChunked, wmb.MlHost)
    assert t2.get_wholememory_handle().get_local_memory()[1] == 64 * 8 * 4
    wmb.destroy_wholememory_tensor(t2)
    wmb.destroy_communicator(comm2)
    print("ALL-OK")
    """)


@pytest.mark.gpu
def test_embedding_handles_after_finalize():
    _run(PRELUDE + """
    comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    emb = wmb.create_embedding(desc(64, 8), comm, wmb.MtChunked, wmb.MlHost, wmb.create_non_cache_policy())
    c_e = emb.wm_embedding
    lib.wholememory_embedding_get_embedding_tensor.restype = ctypes.c_void_p
    assert lib.wholememory_embedding_get_embedding_tensor(c_e) is not None
    wmb.finalize()
    # the embedding object survives, its table does not: gather is refused, destroy releases the host objects
    assert lib.wholememory_embedding_gather(c_e, None, None, False, None, 0) == INVALID_INPUT
    assert lib.wholememory_embedding_gather_gradient_apply(c_e, None, None, False, ctypes.c_float(0.1), None, 0) == INVALID_INPUT
    assert lib.wholememory_destroy_embedding(c_e) == 0
    assert lib.wholememory_destroy_embedding(c_e) == INVALID_INPUT          # twice
    assert lib.wholememory_embedding_get_embedding_tensor(c_e) is None
    assert lib.wholememory_embedding_set_optimizer(c_e, None) == INVALID_INPUT
    o = ctypes.c_void_p()
    assert lib.wholememory_create_embedding_optimizer(ctypes.byref(o), 1) == 0
    lib.wholememory_destroy_embedding_optimizer.restype = None
    lib.wholememory_destroy_embedding_optimizer(o)
    lib.wholememory_destroy_embedding_optimizer(o)                          # twice: ignored, not a double free
    v = ctypes.c_float(0.5)
    assert lib.wholememory_optimizer_set_parameter(o, b"weight_decay", ctypes.byref(v)) == INVALID_INPUT
    print("ALL-OK")
    """)


def test_round1_crash_through_the_reference_binding():
    """The exact sequence of the round-1 GPU run: the reference's cython module (linked to the same libwholegraph.so)
    finalizes the library while this repo's binding still caches a communicator."""
    if not HAVE_REF_BINDING:
        pytest.skip("oracle/_ref/refbinding not built (needs /root/reference + cython at build time)")
    _run(PRELUDE + """
    comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    sys.path.insert(0, %r)
    import wholememory_binding as rwmb
    rwmb.init(0)
    c2 = rwmb.create_communicator(rwmb.create_unique_id(), 0, 1)
    rwmb.destroy_communicator(c2)
    rwmb.finalize()
    refused(wmb.create_embedding, desc(100, 8), comm, wmb.MtChunked, wmb.MlHost, wmb.create_non_cache_policy())
    # what tests/gpu_utils.single_comm() does to recover
    refused(comm.get_rank)
    wmb.init(0, wmb.WholeMemoryLogLevel.LevFatal)
    comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
    assert comm.get_rank() == 0 and comm.get_size() == 1
    print("ALL-OK")
    """ % BIND_DIR)
