"""ctypes view of the C ABI exported by ``wholegraph_b200/lib/libwholegraph.so``.

This is the Python side of the drop-in boundary: struct layouts, enum values and prototypes
mirror ``include/wholememory/*.h`` (which in turn cite the reference headers they replace).
The reference binds the same symbols from cython
(``python/pylibwholegraph/pylibwholegraph/binding/wholememory_binding.pyx:45-216``).

There is no fallback: if the shared library is missing the import fails loudly.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_bool, c_char, c_char_p, c_float, c_int, c_int64, c_size_t,
                    c_uint, c_ulonglong, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwholegraph.so")
# No environment switch here.  The parity tests and the bench run this same binding on the reference's own library by
# executing this file with the path pre-seeded (oracle/ref_lib_loader.py, outside the package); symbols that library
# lacks are tolerated in that mode only.
_ALT = globals().get("_LIB_PATH_PRESET")
if _ALT:
    LIB_PATH = _ALT

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found. Build it with `make -C wholegraph_b200/csrc -j8` "
        "(or __graft_entry__.build()). wholegraph_b200 has no pure-Python / CPU fallback."
    )

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

WHOLEMEMORY_MAX_TENSOR_DIM = 8
WHOLEMEMORY_UNIQUE_ID_BYTES = 128


class TensorDescription(Structure):
    """wholememory_tensor_description_t (include/wholememory/tensor_description.h)"""
    _fields_ = [
        ("sizes", c_int64 * WHOLEMEMORY_MAX_TENSOR_DIM),
        ("strides", c_int64 * WHOLEMEMORY_MAX_TENSOR_DIM),
        ("storage_offset", c_int64),
        ("dim", c_int),
        ("dtype", c_int),
    ]


class UniqueId(Structure):
    _fields_ = [("internal", c_char * WHOLEMEMORY_UNIQUE_ID_BYTES)]


class GlobalReference(Structure):
    """wholememory_gref_t (include/wholememory/global_reference.h)"""
    _fields_ = [
        ("pointer", c_void_p),
        ("rank_memory_offsets", POINTER(c_size_t)),
        ("world_size", c_int),
        ("stride", c_size_t),
        ("same_chunk", c_bool),
    ]


class CliqueInfo(Structure):
    _fields_ = [(n, c_int) for n in ("is_in_clique", "clique_first_rank", "clique_rank", "clique_rank_num",
                                     "clique_id", "clique_num")]


CREATE_CTX_FN = ctypes.CFUNCTYPE(None, POINTER(c_void_p), c_void_p)
DESTROY_CTX_FN = ctypes.CFUNCTYPE(None, c_void_p, c_void_p)
MALLOC_FN = ctypes.CFUNCTYPE(c_void_p, POINTER(TensorDescription), c_int, c_void_p, c_void_p)
FREE_FN = ctypes.CFUNCTYPE(None, c_void_p, c_void_p)


class TempMemoryFns(Structure):
    _fields_ = [
        ("create_memory_context_fn", CREATE_CTX_FN),
        ("destroy_memory_context_fn", DESTROY_CTX_FN),
        ("malloc_fn", MALLOC_FN),
        ("free_fn", FREE_FN),
        ("global_context", c_void_p),
    ]


class OutputMemoryFns(Structure):
    _fields_ = [("malloc_fn", MALLOC_FN), ("free_fn", FREE_FN), ("global_context", c_void_p)]


class EnvFns(Structure):
    """wholememory_env_func_t (include/wholememory/env_func_ptrs.h)"""
    _fields_ = [("temporary_fns", TempMemoryFns), ("output_fns", OutputMemoryFns)]


def _proto(name, restype, *argtypes):
    try:
        fn = getattr(lib, name)  # AttributeError here == symbol missing from the .so: fail loudly
    except AttributeError:
        if _ALT:
            return None
        raise
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


P = POINTER
err = c_int
comm_t = c_void_p
handle_t = c_void_p
tensor_t = c_void_p

# every symbol the reference's cython binding links against (SURVEY 8(b), 64 symbols) + the rest of the headers
_proto("wholememory_init", err, c_uint, c_int)
_proto("wholememory_finalize", err)
_proto("wholememory_create_unique_id", err, P(UniqueId))
_proto("wholememory_create_communicator", err, P(comm_t), UniqueId, c_int, c_int)
_proto("wholememory_split_communicator", err, P(comm_t), comm_t, c_int, c_int)
_proto("wholememory_destroy_communicator", err, comm_t)
_proto("wholememory_communicator_support_type_location", err, comm_t, c_int, c_int)
_proto("wholememory_communicator_get_rank", err, P(c_int), comm_t)
_proto("wholememory_communicator_get_size", err, P(c_int), comm_t)
_proto("wholememory_communicator_get_local_size", err, P(c_int), comm_t)
_proto("wholememory_communicator_get_clique_info", err, P(CliqueInfo), comm_t)
_proto("wholememory_communicator_is_bind_to_nvshmem", c_bool, comm_t)
_proto("wholememory_communicator_set_distributed_backend", err, comm_t, c_int)
_proto("wholememory_communicator_get_distributed_backend", c_int, comm_t)
_proto("wholememory_communicator_barrier", err, comm_t)
_proto("wholememory_is_intranode_communicator", c_bool, comm_t)
_proto("wholememory_is_intra_mnnvl_communicator", c_bool, comm_t)
_proto("wholememory_is_build_with_nvshmem", c_bool)
_proto("wholememory_malloc", err, P(handle_t), c_size_t, comm_t, c_int, c_int, c_size_t, P(c_size_t))
_proto("wholememory_free", err, handle_t)
_proto("wholememory_get_communicator", err, P(comm_t), handle_t)
_proto("wholememory_get_local_communicator", err, P(comm_t), handle_t)
_proto("wholememory_get_cross_communicator", err, P(comm_t), handle_t)
_proto("wholememory_get_memory_type", c_int, handle_t)
_proto("wholememory_get_memory_location", c_int, handle_t)
_proto("wholememory_get_distributed_backend", c_int, handle_t)
_proto("wholememory_get_total_size", c_size_t, handle_t)
_proto("wholememory_get_data_granularity", c_size_t, handle_t)
_proto("wholememory_get_local_memory", err, P(c_void_p), P(c_size_t), P(c_size_t), handle_t)
_proto("wholememory_get_local_size", err, P(c_size_t), handle_t)
_proto("wholememory_get_local_offset", err, P(c_size_t), handle_t)
_proto("wholememory_get_rank_memory", err, P(c_void_p), P(c_size_t), P(c_size_t), c_int, handle_t)
_proto("wholememory_equal_entry_partition_plan", err, P(c_size_t), c_size_t, c_int)
_proto("wholememory_get_global_pointer", err, P(c_void_p), handle_t)
_proto("wholememory_get_global_reference", err, P(GlobalReference), handle_t)
_proto("wholememory_get_rank_partition_sizes", err, P(c_size_t), handle_t)
_proto("wholememory_get_rank_partition_offsets", err, P(c_size_t), handle_t)
_proto("fork_get_device_count", c_int)
_proto("wholememory_load_from_file", err, handle_t, c_size_t, c_size_t, c_size_t, P(c_char_p), c_int, c_int)
_proto("wholememory_store_to_file", err, handle_t, c_size_t, c_size_t, c_size_t, c_char_p)
_proto("wholememory_dtype_get_element_size", c_size_t, c_int)
_proto("wholememory_dtype_is_floating_number", c_bool, c_int)
_proto("wholememory_dtype_is_integer_number", c_bool, c_int)
_proto("wholememory_initialize_tensor_desc", None, P(TensorDescription))
_proto("wholememory_squeeze_tensor", c_bool, P(TensorDescription), c_int)
_proto("wholememory_unsqueeze_tensor", c_bool, P(TensorDescription), c_int)
_proto("wholememory_get_memory_element_count_from_tensor", c_int64, P(TensorDescription))
_proto("wholememory_get_memory_size_from_tensor", c_int64, P(TensorDescription))
_proto("wholememory_create_continuous_global_reference", GlobalReference, c_void_p)
_proto("wholememory_create_tensor", err, P(tensor_t), P(TensorDescription), comm_t, c_int, c_int, P(c_size_t))
_proto("wholememory_destroy_tensor", err, tensor_t)
_proto("wholememory_make_tensor_from_pointer", err, P(tensor_t), c_void_p, P(TensorDescription))
_proto("wholememory_make_tensor_from_handle", err, P(tensor_t), handle_t, P(TensorDescription))
_proto("wholememory_tensor_has_handle", c_bool, tensor_t)
_proto("wholememory_tensor_get_memory_handle", handle_t, tensor_t)
_proto("wholememory_tensor_get_tensor_description", P(TensorDescription), tensor_t)
_proto("wholememory_tensor_get_global_reference", err, tensor_t, P(GlobalReference))
_proto("wholememory_tensor_map_local_tensor", err, tensor_t, P(tensor_t))
_proto("wholememory_tensor_get_data_pointer", c_void_p, tensor_t)
_proto("wholememory_tensor_get_entry_offsets", err, P(c_size_t), tensor_t)
_proto("wholememory_tensor_get_entry_partition_sizes", err, P(c_size_t), tensor_t)
_proto("wholememory_tensor_get_local_entry_count", err, P(c_size_t), tensor_t)
_proto("wholememory_tensor_get_local_entry_start", err, P(c_size_t), tensor_t)
_proto("wholememory_tensor_get_subtensor", err, tensor_t, P(c_int64), P(c_int64), P(tensor_t))
_proto("wholememory_tensor_get_root", tensor_t, tensor_t)
_proto("get_wholememory_tensor_count", c_int64)
_proto("wholememory_gather", err, tensor_t, tensor_t, tensor_t, P(EnvFns), c_void_p, c_int)
_proto("wholememory_scatter", err, tensor_t, tensor_t, tensor_t, P(EnvFns), c_void_p, c_int)
_proto("wholememory_env_test_op", err, tensor_t, tensor_t, c_void_p, c_void_p, c_void_p, c_int64, P(EnvFns), c_void_p)
_proto("wholememory_create_embedding_optimizer", err, P(c_void_p), c_int)
_proto("wholememory_optimizer_set_parameter", err, c_void_p, c_char_p, c_void_p)
_proto("wholememory_destroy_embedding_optimizer", None, c_void_p)
_proto("wholememory_create_embedding_cache_policy", err, P(c_void_p), comm_t, c_int, c_int, c_int, c_float)
_proto("wholememory_destroy_embedding_cache_policy", err, c_void_p)
_proto("wholememory_create_embedding", err, P(c_void_p), P(TensorDescription), comm_t, c_int, c_int, c_void_p,
       P(c_size_t), c_int, c_int)
_proto("wholememory_destroy_embedding", err, c_void_p)
_proto("wholememory_embedding_get_embedding_tensor", tensor_t, c_void_p)
_proto("wholememory_embedding_set_optimizer", err, c_void_p, c_void_p)
_proto("wholememory_embedding_gather", err, c_void_p, tensor_t, tensor_t, c_bool, P(EnvFns), c_int64)
_proto("wholememory_embedding_gather_gradient_apply", err, c_void_p, tensor_t, tensor_t, c_bool, c_float, P(EnvFns),
       c_int64)
_proto("wholememory_embedding_get_optimizer_state_names", P(c_char_p), c_void_p)
_proto("wholememory_embedding_get_optimizer_state", tensor_t, c_void_p, c_char_p)
_proto("wholememory_embedding_writeback_cache", err, c_void_p, c_int64)
_proto("wholememory_embedding_drop_all_cache", err, c_void_p, c_int64)
_proto("wholegraph_csr_unweighted_sample_without_replacement", err, tensor_t, tensor_t, tensor_t, c_int, tensor_t,
       c_void_p, c_void_p, c_void_p, c_ulonglong, P(EnvFns), c_void_p)
_proto("wholegraph_csr_weighted_sample_without_replacement", err, tensor_t, tensor_t, tensor_t, tensor_t, c_int,
       tensor_t, c_void_p, c_void_p, c_void_p, c_ulonglong, P(EnvFns), c_void_p)
_proto("generate_random_positive_int_cpu", err, c_int64, c_int64, tensor_t)
_proto("generate_exponential_distribution_negative_float_cpu", err, c_int64, c_int64, tensor_t)
_proto("graph_append_unique", err, tensor_t, tensor_t, c_void_p, tensor_t, P(EnvFns), c_void_p)
_proto("csr_add_self_loop", err, tensor_t, tensor_t, tensor_t, tensor_t, c_void_p)
_proto("get_device_prop", c_void_p, c_int)
# additions of this build (plain-C access to the built-in env functions)
_proto("wgb200_default_env_func", P(EnvFns))
_proto("wgb200_cached_env_func", P(EnvFns))
_proto("wgb200_drop_cached_env_func_cache", None)

#: symbols declared in include/wholememory/*.h -- tests check that each one is exported
DECLARED_SYMBOLS = sorted(n for n in dir(lib) if False)  # filled below


def _declared():
    import re
    inc = os.path.join(os.path.dirname(_HERE), "include", "wholememory")
    names = set()
    if not os.path.isdir(inc):
        return []
    for f in os.listdir(inc):
        if not f.endswith(".h"):
            continue
        text = open(os.path.join(inc, f)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        text = text.split("namespace wholememory")[0] if "namespace wholememory" in text else text
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
            n = m.group(1)
            if n.startswith(("wholememory_", "wholegraph_", "generate_", "graph_", "csr_", "fork_", "get_")):
                names.add(n)
    # typedef'd callback types are not symbols
    return sorted(n for n in names if not n.endswith("_func_t") and not n.endswith("_fn"))


DECLARED_SYMBOLS = _declared()
