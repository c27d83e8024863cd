"""The torch-level package exposes the reference's pylibwholegraph.torch names for the hot path
(python/pylibwholegraph/pylibwholegraph/torch/__init__.py:14-78, minus the GNN-model / CLI / data-loader glue that
SURVEY.md section 2 row 23 marks out of scope), with the reference's argument names."""
import inspect

import pytest

import wholegraph_b200.torch as wgth

HOT_PATH_NAMES = [
    "WholeMemoryCommunicator", "create_group_communicator", "destroy_communicator", "get_global_communicator",
    "get_local_node_communicator", "get_local_device_communicator", "split_communicator", "get_local_mnnvl_communicator",
    "WholeMemoryOptimizer", "create_wholememory_optimizer", "destroy_wholememory_optimizer",
    "WholeMemoryCachePolicy", "create_builtin_cache_policy", "create_wholememory_cache_policy", "destroy_wholememory_cache_policy",
    "WholeMemoryEmbedding", "create_embedding", "create_embedding_from_filelist", "destroy_embedding", "WholeMemoryEmbeddingModule",
    "init", "init_torch_env", "init_torch_env_and_create_wm_comm", "finalize",
    "WholeMemoryTensor", "create_wholememory_tensor", "create_wholememory_tensor_from_filelist", "destroy_wholememory_tensor",
    "GraphStructure", "get_part_file_name", "get_part_file_list", "compile_cpp_extension",
]


@pytest.mark.parametrize("name", HOT_PATH_NAMES)
def test_name_is_exported(name):
    assert hasattr(wgth, name), name


def test_submodules_match_the_reference_layout():
    from wholegraph_b200.torch import (comm, dlpack_utils, embedding, graph_ops, graph_structure, initialize, tensor, utils,  # noqa: F401
                                       wholegraph_env, wholegraph_ops, wholememory_ops)
    assert callable(wholememory_ops.wholememory_gather_forward_functor)
    assert callable(wholememory_ops.wholememory_scatter_functor)
    assert callable(dlpack_utils.torch_import_from_dlpack)
    assert callable(graph_ops.append_unique) and callable(graph_ops.add_csr_self_loop)


SIGNATURES = {
    "create_wholememory_tensor": ["comm", "memory_type", "memory_location", "sizes", "dtype", "strides", "tensor_entry_partition"],
    "create_embedding": ["comm", "memory_type", "memory_location", "dtype", "sizes", "cache_policy", "embedding_entry_partition",
                         "random_init", "gather_sms", "round_robin_size"],
    "create_embedding_from_filelist": ["comm", "memory_type", "memory_location", "filelist", "dtype", "last_dim_size", "cache_policy",
                                       "embedding_entry_partition", "gather_sms", "round_robin_size"],
    "create_builtin_cache_policy": ["builtin_cache_type", "embedding_memory_type", "embedding_memory_location", "access_type",
                                    "cache_ratio", "cache_memory_type", "cache_memory_location"],
    "create_wholememory_optimizer": ["embeddings", "optimizer_type", "param_dict", "global_comm"],
    "create_wholememory_tensor_from_filelist": ["comm", "memory_type", "memory_location", "filelist", "dtype", "last_dim_size",
                                                "last_dim_strides", "tensor_entry_partition"],
}


@pytest.mark.parametrize("name", sorted(SIGNATURES))
def test_argument_names_follow_the_reference(name):
    assert list(inspect.signature(getattr(wgth, name)).parameters) == SIGNATURES[name]


def test_builtin_cache_policy_argument_checks():
    assert wgth.create_builtin_cache_policy("none", "chunked", "cuda", "readonly", 0.5) is None
    with pytest.raises(ValueError):
        wgth.create_builtin_cache_policy("none", "bogus", "cuda", "readonly", 0.5)
    with pytest.raises(ValueError):
        wgth.create_builtin_cache_policy("none", "chunked", "disk", "readonly", 0.5)
    with pytest.raises(ValueError):
        wgth.create_builtin_cache_policy("everywhere", "chunked", "cuda", "readonly", 0.5)
    with pytest.raises(ValueError):
        wgth.create_builtin_cache_policy("local_device", "chunked", "cuda", "readonly", 0.5, cache_memory_location="disk")


def test_file_entry_counting(tmp_path):
    from wholegraph_b200.torch.utils import count_file_entries, get_part_file_list
    names = get_part_file_list(str(tmp_path / "feat"), 3)
    assert [n.rsplit("/", 1)[1] for n in names] == ["feat_part_0_of_3", "feat_part_1_of_3", "feat_part_2_of_3"]
    for i, n in enumerate(names):
        with open(n, "wb") as f:
            f.write(b"\0" * (64 * (i + 1)))
    assert count_file_entries(names, 64) == 6
    assert count_file_entries(names, 4) == 96
    with pytest.raises(ValueError):
        count_file_entries(names, 48)
    with pytest.raises(ValueError):
        count_file_entries([str(tmp_path / "missing")], 4)
