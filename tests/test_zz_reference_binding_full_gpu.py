"""Drop-in proof beyond gather/scatter: the reference's UNMODIFIED cython binding on this repo's library runs a LazyAdam
training step, embedding gather, unweighted sampling, append_unique and csr_add_self_loop (tests/ref_binding_worker.py,
own process), each checked against the oracle.  Complements tests/test_reference_binding_gpu.py.

(File name sorts last on purpose: added without a GPU at hand.)"""
import glob
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = bool(glob.glob(os.path.join(ROOT, "oracle", "_ref", "refbinding", "wholememory_binding*.so")))


@pytest.mark.skipif(not HAVE, reason="oracle/_ref/refbinding not built (needs /root/reference + cython at build time)")
def test_reference_cython_binding_training_step_sampling_and_graph_ops():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_binding_worker.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "reference binding worker OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
