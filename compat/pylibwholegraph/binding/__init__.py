"""pylibwholegraph.binding.wholememory_binding -> wholegraph_b200.binding (the ctypes binding of this repo, which carries
the names of the reference's cython module)."""
import sys

import wholegraph_b200.binding as _impl

sys.modules[__name__ + ".wholememory_binding"] = _impl
wholememory_binding = _impl
