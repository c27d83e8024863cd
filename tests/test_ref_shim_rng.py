"""The RAFT random stand-in used to build the reference's sampler into oracle/_ref
(oracle/ref_shim/raft/random/rng_device.cuh) must produce exactly the stream the oracle and this repo's sampler use,
otherwise the reference-binary sampling comparison would compare different random numbers.  CPU only: the header is
plain host+device C++, compiled here with g++ into a tiny program.  Also checked: the published pcg32 known-answer
vector (pcg32 demo, seed 42 / stream 54), which pins the core generator -- not RAFT's seeding -- to the PCG paper."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "ref_shim")

PROGRAM = r"""
#include <cstdio>
#include <cstdlib>
#include <raft/random/rng_device.cuh>
int main(int argc, char** argv)
{
  unsigned long long seed = strtoull(argv[1], nullptr, 10), sub = strtoull(argv[2], nullptr, 10);
  int n = atoi(argv[3]);
  raft::random::RngState st(seed, 0, raft::random::GeneratorType::GenPC);
  raft::random::detail::DeviceState<raft::random::detail::PCGenerator> ds(st);
  {
    raft::random::detail::PCGenerator g(ds, sub);
    for (int i = 0; i < n; ++i) {
      raft::random::detail::UniformDistParams<int32_t> p;
      p.start = 0;
      p.end   = 1;
      int32_t v;
      raft::random::detail::custom_next(g, &v, p, 0, 0);
      printf("i32 %d\n", v);
    }
  }
  {
    raft::random::detail::PCGenerator g(ds, sub);
    for (int i = 0; i < n; ++i) {
      uint32_t v;
      g.next(v);
      printf("u32 %u\n", v);
    }
  }
  {
    raft::random::detail::PCGenerator g(ds, sub);
    for (int i = 0; i < n / 2; ++i) {
      int64_t v;
      g.next(v);
      printf("i64 %lld\n", (long long)v);
    }
  }
  return 0;
}
"""


@pytest.fixture(scope="module")
def shim_program(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    src, exe = d / "shim_rng.cpp", d / "shim_rng"
    src.write_text(PROGRAM)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", SHIM, str(src), "-o", str(exe)])
    return str(exe)


def _run(exe, seed, sub, n):
    out = subprocess.run([exe, str(seed), str(sub), str(n)], capture_output=True, text=True, check=True).stdout.split("\n")
    vals = {"i32": [], "u32": [], "i64": []}
    for line in out:
        if line:
            k, v = line.split()
            vals[k].append(int(v))
    return vals


@pytest.mark.parametrize("seed,sub", [(42, 54), (0, 0), (77, 123456789), (2 ** 63 + 5, 2 ** 40 + 3)])
def test_stand_in_stream_equals_the_oracle_stream(shim_program, oracle, seed, sub):
    n = 64
    got = _run(shim_program, seed, sub, n)
    exp = oracle.random_positive_ints(seed, sub, n)
    assert got["i32"] == exp.tolist()
    assert [v & 0x7FFFFFFF for v in got["u32"]] == exp.tolist()
    # int64 draws: low word first, sign bit cleared
    u = got["u32"]
    assert got["i64"] == [((u[2 * i] | (u[2 * i + 1] << 32)) & 0x7FFFFFFFFFFFFFFF) for i in range(n // 2)]


def test_pcg32_known_answer_vector(shim_program):
    """pcg32-demo, `pcg32_srandom(42, 54)`: first six outputs (PCG reference implementation's check file)."""
    got = _run(shim_program, 42, 54, 6)["u32"]
    assert got == [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
