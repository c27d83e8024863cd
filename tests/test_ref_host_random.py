"""The reference's own HOST replay of the sampler's random streams (cpp/src/wholegraph_ops/raft_random_gen.cu, compiled for
the CPU from /root/reference by oracle/build_ref_host_random.sh on top of the restated PCG stand-in) against this repo's
library and the oracle -- on CPU, no GPU needed.

What this pins: everything the reference layers ON TOP of the generator -- the sign-masked int32 / int64 draws the
unweighted sampler consumes, and the weighted sampler's key arithmetic u -> -(0.5 + 0.5u) * 2^-clz -> log1p(u) / log 2
with its "draw again while the 64-bit word is zero" loop (raft_random_gen.cu:27-119).  What it cannot pin: the generator
itself (RAFT is not vendored), which is the same restated PCG on both sides."""
import ctypes
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "ref_host_random.so")

pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/ref_host_random.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def ref(wmb):
    lib = ctypes.CDLL(SO)
    for name in ("ref_generate_random_positive_int_cpu", "ref_generate_exponential_distribution_negative_float_cpu"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
    return lib


CASES = [(42, 54), (0, 0), (1234, 7), (987654321, 2 ** 33 + 11)]


@pytest.mark.parametrize("seed,sub", CASES)
def test_positive_int_streams(ref, wmb, oracle, seed, sub):
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    n = 96
    for dt in (torch.int32, torch.int64):
        theirs, ours = torch.zeros(n, dtype=dt), torch.zeros(n, dtype=dt)
        w = wrap_torch_tensor(theirs)
        assert ref.ref_generate_random_positive_int_cpu(seed, sub, w.get_c_handle()) == 0
        wmb.host_generate_random_positive_int(seed, sub, wrap_torch_tensor(ours))
        assert torch.equal(theirs, ours) and bool((theirs >= 0).all())
    assert theirs.dtype == torch.int64
    exp32 = oracle.random_positive_ints(seed, sub, n)
    t32 = torch.zeros(n, dtype=torch.int32)
    w32 = wrap_torch_tensor(t32)  # keep the wrapper alive across the call: it owns the wholememory_tensor_t
    assert ref.ref_generate_random_positive_int_cpu(seed, sub, w32.get_c_handle()) == 0
    assert t32.tolist() == exp32.tolist()


@pytest.mark.parametrize("seed,sub", CASES)
def test_weighted_key_stream(ref, wmb, oracle, seed, sub):
    from wholegraph_b200.torch.wholegraph_env import wrap_torch_tensor
    n = 4096
    theirs, ours = torch.zeros(n, dtype=torch.float32), torch.zeros(n, dtype=torch.float32)
    w_theirs = wrap_torch_tensor(theirs)  # keep the wrapper alive across the call: it owns the wholememory_tensor_t
    assert ref.ref_generate_exponential_distribution_negative_float_cpu(seed, sub, w_theirs.get_c_handle()) == 0
    wmb.host_generate_exponential_distribution_negative_float(seed, sub, wrap_torch_tensor(ours))
    exp = oracle.exponential_negative_floats(seed, sub, n)
    assert theirs.numpy().tobytes() == ours.numpy().tobytes(), "this repo's host key stream differs from the reference's host code"
    assert theirs.numpy().tobytes() == exp.tobytes(), "the oracle's key stream differs from the reference's host code"
    assert bool((theirs < 0).all()) and bool(torch.isfinite(theirs).all())
