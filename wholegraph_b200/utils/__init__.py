"""Small process helpers (public name of pylibwholegraph/utils)."""
