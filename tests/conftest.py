"""pytest configuration: `gpu` marker, import path, and the oracle build (test infrastructure)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def wmb():
    """The ctypes binding; importing it loads libwholegraph.so (fails loudly when it is not built)."""
    import wholegraph_b200.binding as b
    b.init(0, b.WholeMemoryLogLevel.LevWarn)
    return b


@pytest.fixture(autouse=True)
def _python_env_callbacks_for_host_tests(request):
    """The torch layer allocates op outputs through native (C++) env functions by default.  The host-only tests drive that
    layer over FAKE bindings that call the Python-callback protocol, so they run with the ctypes closures; GPU tests keep
    the default."""
    if request.node.get_closest_marker("gpu") is not None:
        yield
        return
    try:
        import wholegraph_b200.torch.wholegraph_env as wenv
    except Exception:
        yield
        return
    was = wenv.torch_cpp_ext_loaded
    if was:
        wenv.unload_native_env()
    yield
    if was:
        wenv.load_native_env()
