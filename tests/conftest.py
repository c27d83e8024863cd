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
