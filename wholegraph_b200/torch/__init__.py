"""torch-level API, same surface as ``pylibwholegraph.torch`` for the hot path."""
from .comm import (WholeMemoryCommunicator, create_group_communicator, destroy_communicator,
                   get_global_communicator, get_local_device_communicator, get_local_mnnvl_communicator,
                   get_local_node_communicator, set_world_info, split_communicator)
from .initialize import finalize, init, init_torch_env, init_torch_env_and_create_wm_comm
from .tensor import (WholeMemoryTensor, create_wholememory_tensor, create_wholememory_tensor_from_filelist,
                     destroy_wholememory_tensor)
from .embedding import (WholeMemoryCachePolicy, WholeMemoryEmbedding, WholeMemoryEmbeddingModule, WholeMemoryOptimizer,  # noqa: E402
                        create_builtin_cache_policy, create_embedding, create_embedding_from_filelist,
                        create_wholememory_cache_policy, create_wholememory_optimizer,
                        destroy_embedding, destroy_wholememory_cache_policy, destroy_wholememory_optimizer)
from .wholememory_ops import wholememory_gather_forward_functor, wholememory_scatter_functor  # noqa: E402
from .wholegraph_ops import (generate_exponential_distribution_negative_float_cpu, generate_random_positive_int_cpu,  # noqa: E402
                             unweighted_sample_without_replacement, weighted_sample_without_replacement)
from .graph_ops import add_csr_self_loop, append_unique  # noqa: E402
from .graph_structure import GraphStructure  # noqa: E402
from .utils import get_part_file_list, get_part_file_name  # noqa: E402
from .wholegraph_env import compile_cpp_extension  # noqa: E402
