"""pylibwholegraph.test_utils -> wholegraph_b200.test_utils (same module objects)."""
import sys

import wholegraph_b200.test_utils as _impl
import wholegraph_b200.test_utils.test_comm as _tc

sys.modules[__name__ + ".test_comm"] = _tc
sys.modules[__name__] = _impl
