"""Opt-in 256-bit access variant of the row-move kernel (WG_VEC32=1: sm_100's LDG.E.256 / STG.E.256, 32-byte units):
the 13 seeded gather/scatter cases of tests/ref_parity_worker.py -- those whose addresses, strides and row sizes are
multiples of 32 bytes take the new path, the rest fall back -- must stay byte-identical to the oracle.
The knob is read once per process, hence the worker subprocess.

(File name sorts last on purpose: the variant was written without a GPU at hand and is off by default.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vec32_variant_is_byte_identical_to_the_oracle(tmp_path):
    from test_ref_parity_gpu import _oracle_results
    out = str(tmp_path / "vec32.npz")
    env = dict(os.environ, WG_VEC32="1")
    env.pop("WHOLEGRAPH_B200_LIB", None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_parity_worker.py"), out], env=env, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, "worker failed:\n" + p.stdout[-2000:] + p.stderr[-4000:]
    got = np.load(out)
    exp = _oracle_results()
    bad = [k for k, v in exp.items() if not np.array_equal(got[k], v)]
    assert bad == [], "WG_VEC32=1 differs from the oracle on: %s" % bad
