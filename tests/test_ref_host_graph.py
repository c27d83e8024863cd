"""Argument checks of the sampling and graph-op entry points (unweighted / weighted neighbor sampling, graph_append_unique,
csr_add_self_loop) against the REFERENCE SOURCE compiled for the CPU (oracle/_ref/ref_host_graph.so, GPU dispatch targets
stubbed; built by oracle/build_ref_host_graph.sh).  tests/cpp/graph_validation_diff.cpp sends 100,000 operand sets with
wrong ranks, strides and dtypes through both libraries: same error code wherever the reference refuses, "reaches the kernel
launch" wherever it dispatches.  CPU only; on a GPU box the program compares nothing."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_host_graph.so")
OURS = os.path.join(ROOT, "wholegraph_b200", "lib", "libwholegraph.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/ref_host_graph.so not built (needs /root/reference at build time)")
def test_sampling_and_graph_op_argument_checks_equal_the_reference_source(tmp_path):
    exe = str(tmp_path / "graph_validation_diff")
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "tests", "cpp", "graph_validation_diff.cpp"), "-o", exe, "-ldl"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, OURS, REF_SO, "100000"], capture_output=True, text=True, timeout=600)
    lines = [l for l in p.stderr.splitlines() if l.startswith("DIVERGENCE")]
    ok = "100000 iterations, 0 divergences" in p.stdout or "a GPU is present" in p.stdout
    assert p.returncode == 0 and ok, "\n".join(lines[:20]) + "\n" + p.stdout[-500:]
