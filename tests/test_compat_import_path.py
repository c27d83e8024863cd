"""`compat/pylibwholegraph`: code written against the reference's import paths runs on this implementation unchanged."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
import pylibwholegraph.torch as wgth
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.torch.initialize import init_torch_env_and_create_wm_comm
from pylibwholegraph.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
from pylibwholegraph.torch.dlpack_utils import torch_import_from_dlpack
from pylibwholegraph.torch.utils import get_part_file_name
import pylibwholegraph.torch.wholegraph_ops as wg_ops
import pylibwholegraph.torch.graph_ops as graph_ops
import wholegraph_b200.torch, wholegraph_b200.binding, wholegraph_b200.torch.wholegraph_ops
assert wgth is wholegraph_b200.torch and wmb is wholegraph_b200.binding and wg_ops is wholegraph_b200.torch.wholegraph_ops
import pylibwholegraph.torch.comm, wholegraph_b200.torch.comm
assert pylibwholegraph.torch.comm is wholegraph_b200.torch.comm  # one communicator registry, not two
from pylibwholegraph.binding import wholememory_binding as wmb2
assert wmb2 is wmb
assert wgth.WholeMemoryEmbeddingModule and wmb.WholeMemoryMemoryType.MtChunked and wmb.PyWholeMemoryUniqueID
assert get_part_file_name("p", 1, 4) == "p_part_1_of_4"
# a one-rank communicator through the aliased modules: the library really is behind them
wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
c = wmb.create_communicator(wmb.create_unique_id(), 0, 1)
assert c.get_size() == 1
wmb.destroy_communicator(c)
try:
    import pylibwholegraph.torch.gnn_model
    raise SystemExit("out-of-scope module unexpectedly importable")
except ModuleNotFoundError:
    pass
import pylibwholegraph
assert isinstance(pylibwholegraph.__git_commit__, str) and isinstance(pylibwholegraph.__version__, str) and pylibwholegraph.__version__  # the reference's test_version.py
print("compat ok")
"""


def test_reference_import_paths_resolve_to_this_implementation():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat")]))
    p = subprocess.run([sys.executable, "-c", PROGRAM], capture_output=True, text=True, timeout=300, env=env, cwd="/")
    assert p.returncode == 0 and "compat ok" in p.stdout, p.stdout + p.stderr
