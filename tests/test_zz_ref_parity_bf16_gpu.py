"""Three-way bit-exact parity for bf16 (table and/or output side), which the 13 verified cases of test_ref_parity_gpu.py and
the committed golden vectors do not cover: the reference's own gather / scatter kernels (oracle/_ref) vs this repo's
kernels vs the oracle, 9 extra seeded cases incl. double -> bf16 (two rounding hops, like the reference's type_caster).

(File name sorts last on purpose: added without a GPU at hand.)"""
import os
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first execution pending: written after the round-1 GPU budget was spent; remove this marker once it has passed on a B200")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libwholegraph_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libwholegraph_ref.so not built (needs /root/reference at build time)")
def test_bf16_cases_reference_binary_ours_and_oracle(tmp_path, monkeypatch):
    monkeypatch.setenv("WG_PARITY_EXTRA", "1")
    sys.modules.pop("ref_parity_worker", None)
    import test_ref_parity_gpu as T
    try:
        ref = T._run_worker(tmp_path, "ref", REF_SO)
        ours = T._run_worker(tmp_path, "ours")
        exp = T._oracle_results()
        assert len(exp) == 18
        bad = [k for k in exp if not np.array_equal(ref[k], exp[k])]
        assert bad == [], "oracle differs from the reference binary on: %s" % bad
        bad = [k for k in exp if not np.array_equal(ref[k], ours[k])]
        assert bad == [], "this repo's kernels differ from the reference binary on: %s" % bad
    finally:
        sys.modules.pop("ref_parity_worker", None)
