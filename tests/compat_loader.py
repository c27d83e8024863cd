"""Load modules of the reference tree, unchanged, with their `pylibwholegraph...` imports resolved by compat/ (test helper)."""
import importlib
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def activate_compat():
    """Put <repo> and <repo>/compat on sys.path and make sure `pylibwholegraph.binding.wholememory_binding` is THIS repo's
    binding: the reference's compiled cython module (tests/test_binding_surface.py imports it as a top-level module)
    registers itself in sys.modules under that dotted name, which would shadow the alias in the same process."""
    for p in (ROOT, os.path.join(ROOT, "compat")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import wholegraph_b200.binding as ours
    leaf = "pylibwholegraph.binding.wholememory_binding"
    if sys.modules.get(leaf) is not None and sys.modules[leaf] is not ours:
        del sys.modules[leaf]
    import pylibwholegraph.binding.wholememory_binding as aliased
    assert aliased is ours
    return ours


def load_reference_file(path, name, package=None):
    """Execute one .py file of the reference tree as module `name` (not registered in sys.modules).  `package` (e.g.
    "pylibwholegraph.torch") gives the file's relative imports their parent: `from .utils import x` then resolves to the
    aliased module of this repo."""
    activate_compat()
    if package:
        importlib.import_module(package)
        name = package + "." + name
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module
