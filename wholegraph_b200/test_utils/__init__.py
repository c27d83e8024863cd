"""Helpers for tests written against the WholeMemory Python API (public name of pylibwholegraph/test_utils)."""
