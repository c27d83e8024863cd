"""pylibwholegraph.utils -> wholegraph_b200.utils (same module objects)."""
import sys

import wholegraph_b200.utils as _impl
import wholegraph_b200.utils.multiprocess as _mp

sys.modules[__name__ + ".multiprocess"] = _mp
sys.modules[__name__] = _impl
