"""pylibwholegraph.torch and its hot-path submodules -> wholegraph_b200.torch (the same module objects, not copies).

Every submodule is registered under its reference name up front, so `from pylibwholegraph.torch.initialize import ...`
finds the one instance that `wholegraph_b200.torch.initialize` also names (a second copy would carry its own
communicator registry and env-function table)."""
import importlib
import pkgutil
import sys

import wholegraph_b200.torch as _impl

for _m in pkgutil.iter_modules(_impl.__path__):
    sys.modules[__name__ + "." + _m.name] = importlib.import_module("wholegraph_b200.torch." + _m.name)
sys.modules[__name__] = _impl
