"""wholememory::device_reference<float> used inside a kernel on CONTINUOUS, regular-CHUNKED and irregular-CHUNKED global
references handed out by the library (tests/cpp/device_reference_test.cu), 1 rank and 3 ranks sharing the GPU.
The program is compiled here, at build time of the test, with nvcc for sm_100a against include/wholememory only."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "wholegraph_b200", "lib")
EXE = os.path.join(LIB_DIR, "device_reference_test")


def build():
    cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "device_reference_test.cu"), "-o", EXE, "-L", LIB_DIR, "-lwholegraph", "-Xlinker", "-rpath," + LIB_DIR]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]


@pytest.mark.parametrize("ranks", [1, 3])
def test_device_reference_in_a_kernel(ranks):
    if not os.path.exists(EXE):
        build()
    env = dict(os.environ, WG_BOOTSTRAP_TIMEOUT_S="120")
    p = subprocess.run([EXE, str(ranks)], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "0 failed checks" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
