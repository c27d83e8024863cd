"""TEST / BENCH infrastructure only: run this repo's ctypes binding on ANOTHER build of the same C ABI -- the reference's own
library rebuilt under oracle/_ref -- so both can be compared and timed by one harness.

The shipped package has no such switch (wholegraph_b200/_lib.py always loads wholegraph_b200/lib/libwholegraph.so).  The
swap happens here, outside the package: `_lib.py` is executed as the module `wholegraph_b200._lib` with the library
path pre-seeded, BEFORE anything imports the package.  Workers started by the parity tests / bench receive the path in
the environment variable WHOLEGRAPH_B200_LIB and call apply_env() first thing."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def use_library(path):
    if "wholegraph_b200._lib" in sys.modules:
        loaded = sys.modules["wholegraph_b200._lib"].LIB_PATH
        if os.path.abspath(loaded) != os.path.abspath(path):
            raise RuntimeError("wholegraph_b200 is already bound to %s; use_library() must run before the first import" % loaded)
        return
    # torch first: its bundled NCCL (2.28) must be the libnccl.so.2 of the process -- the reference library links the
    # system's older one and, loaded RTLD_GLOBAL before torch, would leave libtorch_cuda.so with unresolved symbols
    try:
        import torch  # noqa: F401
    except ImportError:
        pass
    src = os.path.join(_ROOT, "wholegraph_b200", "_lib.py")
    spec = importlib.util.spec_from_file_location("wholegraph_b200._lib", src)
    mod = importlib.util.module_from_spec(spec)
    mod.__dict__["_LIB_PATH_PRESET"] = os.path.abspath(path)  # read by _lib.py when it executes
    sys.modules["wholegraph_b200._lib"] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules["wholegraph_b200._lib"]
        raise


def apply_env():
    """Honour WHOLEGRAPH_B200_LIB (set by the tests / bench for their worker processes); no-op when unset."""
    path = os.environ.get("WHOLEGRAPH_B200_LIB")
    if path:
        if _ROOT not in sys.path:
            sys.path.insert(0, _ROOT)
        use_library(path if os.path.isabs(path) else os.path.join(_ROOT, path))
