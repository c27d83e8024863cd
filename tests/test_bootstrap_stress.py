"""Control-plane bootstrap under load (tests/cpp/bootstrap_stress.cpp): forked ranks, thousands of collectives on both the
shared-memory mailbox and the socket path, every byte checked; intruders on the abstract socket are dropped."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "wholegraph_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


@pytest.fixture(scope="module")
def program(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("boot") / "bootstrap_stress")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-I", os.path.join(ROOT, "wholegraph_b200", "csrc"), os.path.join(ROOT, "tests", "cpp", "bootstrap_stress.cpp"), "-o", exe,
           "-L", LIB_DIR, "-lwholegraph", "-Wl,-rpath," + LIB_DIR, "-L", os.path.join(CUDA, "lib64"), "-lcudart"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    return exe


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_collectives_on_mailbox_and_sockets(program, ranks):
    env = dict(os.environ, WG_BOOTSTRAP_TIMEOUT_S="120")
    p = subprocess.run([program, str(ranks), "3000"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "0 failed ranks" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_strangers_on_the_socket_are_dropped(program):
    """One intruder sends a plausible rank with a wrong secret, one connects and stays silent: the communicator still
    forms with exactly its own ranks."""
    env = dict(os.environ, WG_BOOTSTRAP_TIMEOUT_S="120")
    p = subprocess.run([program, "3", "200", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "0 failed ranks" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
    assert "dropped a connection" in p.stdout + p.stderr
