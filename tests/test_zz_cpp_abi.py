"""C++ callers: tests/cpp/abi_cpp_test.cpp is compiled with plain g++ against include/wholememory/*.h and linked to the
in-tree libwholegraph.so, like the reference's own C++ tests and bench link wholegraph::wholegraph.
CPU mode (here): descriptor helpers, pointer views, C++-only env helpers, single-rank and forked 2-/3-rank communicators,
and device entry points failing loudly without a GPU.  GPU mode: gather / scatter / SGD step / sampling with closed-form
checks, 1 rank and 2 forked ranks sharing the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "wholegraph_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cppabi") / "abi_cpp_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cpp", "abi_cpp_test.cpp"), "-o", exe, "-L", LIB_DIR, "-lwholegraph", "-Wl,-rpath," + LIB_DIR,
           "-L", os.path.join(CUDA, "lib64"), "-lcudart"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    assert "warning" not in p.stderr, p.stderr[-4000:]
    return exe


def test_cpp_caller_host_side(program):
    p = subprocess.run([program, "cpu"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "0 failed checks" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 2])
def test_cpp_caller_on_the_gpu(program, ranks):
    env = dict(os.environ, WG_BOOTSTRAP_TIMEOUT_S="120")  # a diverged rank must end in an error well before the pytest timeout
    p = subprocess.run([program, "gpu", str(ranks)], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "0 failed checks" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
