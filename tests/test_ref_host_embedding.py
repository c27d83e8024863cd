"""Host entry points of the embedding layer against the REFERENCE SOURCE compiled for the CPU
(oracle/_ref/ref_host_embedding.so = the reference's embedding.cpp + embedding_optimizer.cpp + embedding_cache.cpp, built
by oracle/build_ref_host_embedding.sh; GPU-side symbols stay unbound, only calls that never reach the device are made).
Same arguments into both libraries, same return codes expected:
  * wholememory_create_embedding_optimizer for every optimizer type value, valid or not;
  * wholememory_optimizer_set_parameter for every (optimizer, parameter name) pair -- which names each optimizer accepts
    (reference embedding_optimizer.cpp:119, :180-189, :303, :404-409) -- and unknown names;
  * wholememory_create_embedding_cache_policy over a grid of memory types / locations / access types / ratios.
CPU only."""
import ctypes
import itertools
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_host_embedding.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/ref_host_embedding.so not built (needs /root/reference at build time)")

NAMES = ["weight_decay", "epsilon", "beta1", "beta2", "adam_w", "alpha", "lr", "momentum", "", "Beta1"]


@pytest.fixture(scope="module")
def libs(wmb):
    from wholegraph_b200 import _lib
    ref = ctypes.CDLL(REF_SO, mode=os.RTLD_LAZY)
    for lib in (ref, _lib.lib):
        lib.wholememory_create_embedding_optimizer.restype = ctypes.c_int
        lib.wholememory_create_embedding_optimizer.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        lib.wholememory_optimizer_set_parameter.restype = ctypes.c_int
        lib.wholememory_optimizer_set_parameter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        lib.wholememory_destroy_embedding_optimizer.restype = None
        lib.wholememory_destroy_embedding_optimizer.argtypes = [ctypes.c_void_p]
        lib.wholememory_create_embedding_cache_policy.restype = ctypes.c_int
        lib.wholememory_create_embedding_cache_policy.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                                  ctypes.c_int, ctypes.c_float]
        lib.wholememory_destroy_embedding_cache_policy.restype = ctypes.c_int
        lib.wholememory_destroy_embedding_cache_policy.argtypes = [ctypes.c_void_p]
    return {"reference": ref, "ours": _lib.lib}


def _optimizer_matrix(lib):
    rows = {}
    for opt_type in range(-1, 7):
        handle = ctypes.c_void_p()
        rc = lib.wholememory_create_embedding_optimizer(ctypes.byref(handle), opt_type)
        codes = []
        if rc == 0:
            for name in NAMES:
                value = ctypes.c_float(0.25)
                codes.append(lib.wholememory_optimizer_set_parameter(handle, name.encode(), ctypes.byref(value)))
            lib.wholememory_destroy_embedding_optimizer(handle)
        rows[opt_type] = (rc, codes)
    return rows


def test_optimizer_creation_and_parameter_names(libs, capfd):
    ref, ours = _optimizer_matrix(libs["reference"]), _optimizer_matrix(libs["ours"])
    capfd.readouterr()  # both libraries log every refused name
    assert ours == ref
    assert [t for t, (rc, _) in ref.items() if rc == 0] == [1, 2, 3, 4]  # SGD, LazyAdam, RMSProp, AdaGrad


def test_cache_policy_argument_checks(libs, capfd):
    ratios = [0.0, 1.0 / 1024, 1.0 / 512, 0.001953125, 0.25, 1.0, 1.0000001, 2.0, -0.5]
    for mt, loc, acc, ratio in itertools.product(range(0, 5), range(0, 3), range(0, 3), ratios):
        codes = {}
        for who, lib in libs.items():
            handle = ctypes.c_void_p()
            rc = lib.wholememory_create_embedding_cache_policy(ctypes.byref(handle), None, mt, loc, acc, ratio)
            if rc == 0:
                assert lib.wholememory_destroy_embedding_cache_policy(handle) == 0
            codes[who] = rc
        assert codes["ours"] == codes["reference"], (mt, loc, acc, ratio, codes)
    capfd.readouterr()
