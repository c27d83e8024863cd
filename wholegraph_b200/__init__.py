"""wholegraph_b200: B200-native WholeMemory gather / scatter / sparse-optimizer / CSR-sampling engine.

``wholegraph_b200.binding`` mirrors the reference's cython binding over the C ABI in
``lib/libwholegraph.so``; ``wholegraph_b200.torch`` mirrors ``pylibwholegraph.torch``.
Importing the package loads the CUDA shared library and fails if it is missing (no CPU fallback).
"""
from . import _lib  # noqa: F401  (loads libwholegraph.so, raises ImportError when absent)
from . import binding  # noqa: F401

__all__ = ["binding"]
