"""Tensor <-> partition arithmetic of HANDLE-backed tensors against the REFERENCE SOURCE compiled for the CPU
(oracle/_ref/ref_host_tensor.so).  A WholeMemory handle cannot exist without a GPU, so tests/cpp/wm_tensor_diff.cpp builds
synthetic handles (world size 1..8, any rank, equal and custom row partitions, all memory types, padded strides) from this
repo's internal structs; the reference code reads them through the C-ABI accessors, this repo's code directly.  100,000
cases: entry offsets, entry partition sizes, local entry count / start, sub-tensor views, local-tensor mapping, data
pointers.  Zero divergences allowed; the one documented difference (a row-truncated view that ends before this rank's
partition: the reference's unsigned subtraction wraps) is counted separately.  CPU only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_host_tensor.so")
LIB_DIR = os.path.join(ROOT, "wholegraph_b200", "lib")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/ref_host_tensor.so not built (needs /root/reference at build time)")
def test_partition_arithmetic_of_handle_backed_tensors_equals_the_reference_source(tmp_path):
    exe = str(tmp_path / "wm_tensor_diff")
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "wholegraph_b200", "csrc"), "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "tests", "cpp", "wm_tensor_diff.cpp"), "-o", exe, "-L", LIB_DIR, "-lwholegraph",
                        "-Wl,-rpath," + LIB_DIR, "-ldl"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, REF_SO, "100000"], capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stderr.splitlines() if l.startswith("DIVERGENCE")]
    assert p.returncode == 0 and "100000 iterations, 0 divergences" in p.stdout, "\n".join(lines[:20]) + "\n" + p.stdout[-500:]
