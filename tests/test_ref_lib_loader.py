"""oracle/ref_lib_loader.py: the ONLY way to run the binding on another build of the C ABI lives outside the package.
The package itself ignores WHOLEGRAPH_B200_LIB; the loader pre-seeds wholegraph_b200._lib before the first import and
refuses to swap libraries afterwards."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "wholegraph_b200", "lib", "libwholegraph.so")


def _child(body, env=None):
    src = "import sys\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body)
    p = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return p.stdout


def test_the_package_ignores_the_environment_variable():
    out = _child("""
        import wholegraph_b200
        from wholegraph_b200 import _lib
        print(_lib.LIB_PATH)
    """, env={"WHOLEGRAPH_B200_LIB": "/nonexistent/libother.so"})
    assert out.strip().endswith("wholegraph_b200/lib/libwholegraph.so")


def test_loader_preseeds_the_library_and_refuses_a_late_swap(tmp_path):
    other = str(tmp_path / "libwholegraph_copy.so")  # "another build of the same ABI": a byte copy under another path
    with open(OURS, "rb") as f, open(other, "wb") as g:
        g.write(f.read())
    out = _child("""
        from oracle.ref_lib_loader import apply_env, use_library
        apply_env()
        import wholegraph_b200.binding as wmb
        from wholegraph_b200 import _lib
        print(_lib.LIB_PATH)
        wmb.init(0, wmb.WholeMemoryLogLevel.LevFatal)
        comm = wmb.create_communicator(wmb.create_unique_id(), 0, 1)   # the copy answers through the same binding
        assert comm.get_size() == 1
        use_library(_lib.LIB_PATH)                                     # same library again: fine
        try:
            use_library(%r)
            print("SWAPPED")
        except RuntimeError as e:
            print("refused:", "already bound" in str(e))
    """ % OURS, env={"WHOLEGRAPH_B200_LIB": other})
    lines = out.strip().splitlines()
    assert lines[0] == other and lines[-1] == "refused: True", out
