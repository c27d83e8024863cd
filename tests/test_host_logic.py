"""CPU tests of the C-ABI library's host side: it must LOAD without a GPU, export every declared
symbol, keep the reference's descriptor / partition / error-code behaviour, and refuse (loudly)
to do device work when there is no CUDA device -- there is no CPU fallback."""
import ctypes
from ctypes import byref, c_int64, c_size_t, c_void_p

import pytest

from wholegraph_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol():
    assert len(_lib.DECLARED_SYMBOLS) >= 95
    missing = [s for s in _lib.DECLARED_SYMBOLS if not hasattr(_lib.lib, s)]
    assert missing == []


def test_the_64_symbols_the_reference_binding_links_are_present():
    # `nm -u` of the reference's cythonized binding
    needed = """wholememory_init wholememory_finalize wholememory_create_unique_id wholememory_create_communicator
    wholememory_split_communicator wholememory_destroy_communicator wholememory_communicator_support_type_location
    wholememory_communicator_get_rank wholememory_communicator_get_size wholememory_communicator_get_clique_info
    wholememory_communicator_barrier wholememory_communicator_set_distributed_backend
    wholememory_communicator_get_distributed_backend wholememory_is_intranode_communicator wholememory_malloc
    wholememory_free wholememory_get_communicator wholememory_get_local_communicator wholememory_get_cross_communicator
    wholememory_get_memory_type wholememory_get_memory_location wholememory_get_total_size wholememory_get_local_memory
    wholememory_get_rank_memory wholememory_get_global_pointer wholememory_equal_entry_partition_plan
    wholememory_load_from_file wholememory_store_to_file fork_get_device_count wholememory_dtype_get_element_size
    wholememory_create_tensor wholememory_destroy_tensor wholememory_make_tensor_from_pointer
    wholememory_make_tensor_from_handle wholememory_tensor_get_memory_handle wholememory_tensor_get_tensor_description
    wholememory_tensor_get_subtensor wholememory_tensor_get_local_entry_count wholememory_tensor_get_local_entry_start
    get_wholememory_tensor_count wholememory_gather wholememory_scatter wholememory_env_test_op
    wholememory_create_embedding_optimizer wholememory_optimizer_set_parameter wholememory_destroy_embedding_optimizer
    wholememory_create_embedding_cache_policy wholememory_destroy_embedding_cache_policy wholememory_create_embedding
    wholememory_destroy_embedding wholememory_embedding_get_embedding_tensor wholememory_embedding_set_optimizer
    wholememory_embedding_gather wholememory_embedding_gather_gradient_apply
    wholememory_embedding_get_optimizer_state_names wholememory_embedding_get_optimizer_state
    wholememory_embedding_writeback_cache wholememory_embedding_drop_all_cache
    wholegraph_csr_unweighted_sample_without_replacement wholegraph_csr_weighted_sample_without_replacement
    generate_random_positive_int_cpu generate_exponential_distribution_negative_float_cpu graph_append_unique
    csr_add_self_loop""".split()
    assert len(needed) == 64
    for s in needed:
        assert hasattr(_lib.lib, s), s


def test_struct_layouts_match_the_abi():
    assert ctypes.sizeof(_lib.TensorDescription) == 8 * 8 * 2 + 8 + 4 + 4  # sizes, strides, offset, dim, dtype
    assert ctypes.sizeof(_lib.GlobalReference) == 40
    assert ctypes.sizeof(_lib.UniqueId) == 128
    assert ctypes.sizeof(_lib.EnvFns) == 5 * 8 + 3 * 8


def test_dtype_helpers(wmb):
    L = _lib.lib
    sizes = {wmb.DtFloat: 4, wmb.DtHalf: 2, wmb.DtDouble: 8, wmb.DtBF16: 2, wmb.DtInt: 4, wmb.DtInt64: 8, wmb.DtInt16: 2,
             wmb.DtInt8: 1, wmb.DtUnknown: 0}
    for dt, s in sizes.items():
        assert L.wholememory_dtype_get_element_size(int(dt)) == s
    assert L.wholememory_dtype_is_floating_number(int(wmb.DtBF16)) and not L.wholememory_dtype_is_floating_number(int(wmb.DtInt))
    assert L.wholememory_dtype_is_integer_number(int(wmb.DtInt8)) and not L.wholememory_dtype_is_integer_number(int(wmb.DtHalf))


def test_squeeze_unsqueeze(wmb):
    L = _lib.lib
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((10,))
    d.set_stride((1,))
    assert L.wholememory_unsqueeze_tensor(byref(d.tensor_description), 1)
    assert d.shape == (10, 1) and d.stride() == (1, 1)
    assert L.wholememory_squeeze_tensor(byref(d.tensor_description), 1)
    assert d.shape == (10,)
    d.set_shape((4, 1, 6))
    d.set_stride((12, 6, 1))
    assert not L.wholememory_squeeze_tensor(byref(d.tensor_description), 0)  # size != 1
    assert L.wholememory_squeeze_tensor(byref(d.tensor_description), 1) is False  # stride 6 != stride 1 of next dim
    assert not L.wholememory_unsqueeze_tensor(byref(d.tensor_description), 5)
    d.set_shape((3, 5))
    d.set_stride((8, 1))
    assert L.wholememory_get_memory_element_count_from_tensor(byref(d.tensor_description)) == 24
    assert L.wholememory_get_memory_size_from_tensor(byref(d.tensor_description)) == 96


def test_equal_partition_plan(wmb, oracle):
    for n, ws in [(10, 4), (3, 8), (1_000_000_000, 8), (0, 2), (7, 7)]:
        per = wmb.equal_partition_plan(n, ws)
        off = oracle.partition(n, ws)
        assert per == (off[1] - off[0] if n else 0) or n < ws


def test_pointer_tensor_views_and_subtensor_math(wmb):
    before = wmb.py_get_wholememory_tensor_count()
    buf = (ctypes.c_float * 400)()
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((20, 16))
    d.set_stride((20, 1))
    t = wmb.make_tensor_as_wholememory(d, ctypes.addressof(buf))
    assert t.shape == (20, 16) and t.stride() == (20, 1) and t.dim() == 2 and t.dtype == wmb.DtFloat
    assert wmb.py_get_wholememory_tensor_count() == before + 1
    sub = t.get_sub_tensor([2, 3], [7, -1])
    assert sub.shape == (5, 13) and sub.storage_offset() == 2 * 20 + 3 and sub.stride() == (20, 1)
    sub2 = sub.get_sub_tensor([-1, 1], [2, 4])
    assert sub2.shape == (2, 3) and sub2.storage_offset() == 2 * 20 + 3 + 1
    ptr = _lib.lib.wholememory_tensor_get_data_pointer(sub.wholememory_tensor)
    assert ptr == ctypes.addressof(buf) + 4 * (2 * 20 + 3)
    for bad in ([5, 0], [5, 4]), ([0, 16], [1, 17]), ([0, 0], [0, 4]):
        with pytest.raises(ValueError):
            t.get_sub_tensor(*bad)
    assert _lib.lib.wholememory_tensor_get_root(sub2.wholememory_tensor) == t.get_c_handle()
    assert not _lib.lib.wholememory_tensor_has_handle(t.wholememory_tensor)
    assert t.get_local_entry_count() == 20 and t.get_local_entry_start() == 0
    for x in (sub2, sub, t):
        wmb.destroy_wholememory_tensor(x)
    assert wmb.py_get_wholememory_tensor_count() == before


def test_bad_descriptions_are_rejected(wmb):
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((4, 4))
    d.set_stride((4, 2))  # inner stride must be 1
    out = c_void_p()
    buf = (ctypes.c_float * 16)()
    rc = _lib.lib.wholememory_make_tensor_from_pointer(byref(out), ctypes.addressof(buf), byref(d.tensor_description))
    assert rc == wmb.WholeMemoryErrorCode.InvalidInput


def test_single_rank_communicator_and_no_cpu_fallback(wmb):
    """Control plane works without CUDA; anything that needs the device fails with an error code."""
    import torch
    uid = wmb.create_unique_id()
    comm = wmb.create_communicator(uid, 0, 1)
    assert comm.get_rank() == 0 and comm.get_size() == 1
    comm.barrier() if torch.cuda.is_available() else None
    assert comm.get_distributed_backend() == wmb.DbNCCL
    with pytest.raises(NotImplementedError):
        comm.set_distributed_backend(wmb.DbNVSHMEM)
    assert comm.get_clique_info()[0] is False
    assert comm.support_type_location(wmb.MtDistributed, wmb.MlDevice)
    assert not comm.support_type_location(wmb.MtHierarchy, wmb.MlDevice)
    if not torch.cuda.is_available():
        assert not comm.support_type_location(wmb.MtContinuous, wmb.MlDevice)
        for mt, ml in [(wmb.MtDistributed, wmb.MlDevice), (wmb.MtContinuous, wmb.MlHost)]:
            with pytest.raises(RuntimeError):  # CUDAError -> RuntimeError, never a silent host path
                wmb.malloc(1 << 20, comm, mt, ml, 64)
        # invalid arguments are still diagnosed before any device work
        with pytest.raises(ValueError):
            wmb.malloc(100, comm, wmb.MtDistributed, wmb.MlDevice, 64)  # size % granularity != 0
    wmb.destroy_communicator(comm)


def test_gather_without_gpu_fails_loudly(wmb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    from wholegraph_b200.torch.wholegraph_env import get_wholegraph_env_fns, wrap_torch_tensor
    table = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((10, 4))
    d.set_stride((4, 1))
    t = wmb.make_tensor_as_wholememory(d, table.data_ptr())
    idx = torch.tensor([1, 2], dtype=torch.int64)
    out = torch.zeros(2, 4)
    with pytest.raises(RuntimeError):
        wmb.wholememory_gather_op(t, wrap_torch_tensor(idx), wrap_torch_tensor(out), get_wholegraph_env_fns(), 0)
    assert out.abs().sum() == 0  # nothing was computed on the host
    wmb.destroy_wholememory_tensor(t)


def test_optimizer_parameter_names(wmb):
    opt = wmb.create_optimizer(wmb.OptLazyAdam, {"weight_decay": 0.1, "epsilon": 1e-6, "beta1": 0.8, "beta2": 0.9, "adam_w": 1.0})
    opt.destroy_optimizer()
    with pytest.raises(ValueError):
        wmb.create_optimizer(wmb.OptSgd, {"beta1": 0.5})
    with pytest.raises(ValueError):
        wmb.create_optimizer(wmb.OptAdaGrad, {"alpha": 0.5})
    wmb.create_optimizer(wmb.OptRmsProp, {"alpha": 0.5, "epsilon": 1e-3}).destroy_optimizer()


def test_out_of_scope_entry_points_say_so(wmb):
    L = _lib.lib
    # every op diagnoses null arguments instead of crashing
    assert L.wholegraph_csr_weighted_sample_without_replacement(None, None, None, None, 5, None, None, None, None, 0, None,
                                                                None) == wmb.WholeMemoryErrorCode.InvalidInput
    assert L.generate_exponential_distribution_negative_float_cpu(0, 0, None) == wmb.WholeMemoryErrorCode.InvalidInput
    assert L.wholememory_load_from_file(None, 0, 0, 0, None, 0, 0) == wmb.WholeMemoryErrorCode.InvalidInput
    assert L.graph_append_unique(None, None, None, None, None, None) == wmb.WholeMemoryErrorCode.InvalidInput
    assert L.csr_add_self_loop(None, None, None, None, None) == wmb.WholeMemoryErrorCode.InvalidInput
    with pytest.raises(ValueError):
        wmb.create_cache_policy(wmb.PyWholeMemoryComm(None), wmb.MtChunked, wmb.MlDevice, wmb.AtReadOnly, 2.0)


def test_host_key_stream_matches_oracle(wmb, oracle):
    """generate_exponential_distribution_negative_float_cpu: log2(u) keys of weight 1, all negative, oracle-identical."""
    import torch
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    out = torch.zeros(64, dtype=torch.float32)
    wmb.host_generate_exponential_distribution_negative_float(1234, 7, wrap_torch_tensor(out))
    exp = oracle.exponential_negative_floats(1234, 7, 64)
    assert out.numpy().tobytes() == exp.tobytes()
    assert (out < 0).all() and torch.isfinite(out).all()


def test_host_random_stream_matches_oracle(wmb, oracle):
    import torch
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    out = torch.zeros(16, dtype=torch.int32)
    wmb.host_generate_random_positive_int(42, 54, wrap_torch_tensor(out))
    assert out.tolist() == oracle.random_positive_ints(42, 54, 16).tolist()
    out64 = torch.zeros(4, dtype=torch.int64)
    wmb.host_generate_random_positive_int(42, 54, wrap_torch_tensor(out64))
    r = oracle.random_positive_ints(42, 54, 8)  # sign-masked u32 draws; rebuild the u64 pairs from raw values
    assert all(v >= 0 for v in out64.tolist())
