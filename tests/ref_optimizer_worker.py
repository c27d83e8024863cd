"""Runs a fixed, seeded list of sparse-optimizer cases and dumps the resulting weights / optimizer states.

Three legs, same inputs:
  * WHOLEGRAPH_B200_LIB unset  -> this repo's library through its public path
    (wholememory_embedding_gather_gradient_apply: fused duplicate merge + update kernel);
  * WHOLEGRAPH_B200_LIB = oracle/_ref/libwholegraph_ref.so, WG_OPT_WORKER_MODE=api -> the REFERENCE's whole
    gather_gradient_apply pipeline (cpp/src/wholememory/embedding.cpp:146-323, built from /root/reference) through the
    same public calls;
  * WHOLEGRAPH_B200_LIB = oracle/_ref/libwholegraph_ref.so (default mode "hook") -> the REFERENCE's own dedup + optimizer kernels
    (cpp/src/wholememory_ops/functions/exchange_embeddings_nccl_func.cu:76-206, embedding_optimizer_func.cu) through
    the test hook oracle/ref_optimizer_hook.cpp, on plain device buffers laid out like the reference's embedding
    (row stride padded to 4 floats, LazyAdam state [N, 2*stride] = [m | v], per-row [beta1^t, beta2^t] starting at 1).
Used by test_zz_ref_optimizer_parity_gpu.py."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_lib_loader import apply_env  # noqa: E402  (WHOLEGRAPH_B200_LIB: run this harness on the reference's library)

apply_env()
sys.path.insert(0, os.path.join(ROOT, "tests"))

SMALL = os.environ.get("WG_GOLDEN_SMALL") == "1"  # compact sizes for the committed golden vectors (tools/make_golden.sh)
ROWS = 100 if SMALL else 3000
STEPS = 3
LR = 0.05
# (kind, params, dim, index dtype, gradients per step)
CASES = [
    ("sgd", {"weight_decay": 0.02}, 128, np.int64, 2500),
    ("adam", {}, 127, np.int64, 2500),
    ("adam", {"adam_w": 1.0, "weight_decay": 0.01, "beta1": 0.8}, 392, np.int32, 2500),
    ("adagrad", {"epsilon": 1e-6}, 132, np.int64, 2500),
    ("rmsprop", {"alpha": 0.9, "weight_decay": 0.001}, 513, np.int32, 1500),
    ("adam", {}, 1024, np.int64, 1000),
    ("adam", {}, 4, np.int64, 3000),
    ("sgd", {}, 1, np.int32, 3000),
]
DEFAULTS = {"weight_decay": 0.0, "epsilon": 1e-8, "beta1": 0.9, "beta2": 0.999, "adam_w": 0.0, "alpha": 0.99}
OPT_ID = {"sgd": 1, "adam": 2, "rmsprop": 3, "adagrad": 4}  # wholememory_optimizer_type_t


def case_inputs(ci):
    """Initial weights and the per-step (indices, gradients): valid ids only, heavy duplication (Zipf)."""
    kind, params, dim, idt, n = CASES[ci]
    n = max(40, n // 25) if SMALL else n
    rng = np.random.default_rng(4242 + ci)
    w0 = rng.standard_normal((ROWS, dim)).astype(np.float32)
    steps = []
    for _ in range(STEPS):
        idx = (rng.zipf(1.2, size=n) % ROWS).astype(idt)
        g = rng.standard_normal((n, dim)).astype(np.float32)
        steps.append((idx, g))
    return w0, steps


def run_api(ci, with_states=True):
    """Through the public C ABI (create_embedding / set_optimizer / gather_gradient_apply) of whichever library is loaded."""
    import torch
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    kind, params, dim, idt, n = CASES[ci]
    comm = G.single_comm()
    w0, steps = case_inputs(ci)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((ROWS, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, wmb.MtChunked, wmb.MlDevice, wmb.create_non_cache_policy())
    opt = wmb.create_optimizer({"sgd": wmb.OptSgd, "adam": wmb.OptLazyAdam, "adagrad": wmb.OptAdaGrad, "rmsprop": wmb.OptRmsProp}[kind], params)
    opt.add_embedding(emb)
    local, off = emb.get_embedding_tensor().get_local_tensor(wmb.MlDevice, 0)
    local.copy_(torch.from_numpy(w0))
    env = get_wholegraph_env_fns()
    for idx, g in steps:
        wmb.EmbeddingGatherGradientApply(emb, wrap_torch_tensor(torch.from_numpy(idx).cuda()), wrap_torch_tensor(torch.from_numpy(g).cuda()),
                                         False, LR, env, get_stream())
    torch.cuda.synchronize()
    res = {"w": local.cpu().numpy().copy()}
    # the reference's get_optimizer_state views cover rows [0, D) only (embedding.cpp:336), so states are read back from this repo's library only
    names = {"adam": ["m", "v", "beta12t"], "adagrad": ["state_sum"], "rmsprop": ["v"], "sgd": []}[kind] if with_states else []
    for nm in names:
        res[nm] = emb.get_optimizer_state(nm).get_local_tensor(wmb.MlDevice, 0)[0].cpu().numpy()[:, : (2 if nm == "beta12t" else dim)].copy()
    emb.destroy_embedding()
    opt.destroy_optimizer()
    return res


_INITED = False


def _init_once(wmb):
    """The reference's wholememory_init refuses a second call (initialize.cpp:40); this repo's is idempotent."""
    global _INITED
    if not _INITED:
        wmb.init(0, wmb.WholeMemoryLogLevel.LevWarn)
        _INITED = True


def run_reference(ci):
    import torch
    import wholegraph_b200.binding as wmb
    from wholegraph_b200 import _lib
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    kind, params, dim, idt, n = CASES[ci]
    p = dict(DEFAULTS)
    p.update(params)
    fn = _lib.lib.wgref_dedup_and_optimizer_step
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int64] + [ctypes.c_float] * 4 + [ctypes.c_int] + \
                  [ctypes.c_float] * 2 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    torch.cuda.set_device(0)
    _init_once(wmb)
    w0, steps = case_inputs(ci)
    stride = (dim + 3) // 4 * 4  # align_embedding_dim for fp32: 16-byte rows (reference embedding.cpp:43-50)
    w = torch.zeros(ROWS, stride, device="cuda")
    w[:, :dim].copy_(torch.from_numpy(w0))
    per_elem = {"adam": 2, "adagrad": 1, "rmsprop": 1, "sgd": 0}[kind]
    state = torch.zeros(ROWS, per_elem * stride, device="cuda") if per_elem else None
    b12 = torch.ones(ROWS, 2, device="cuda") if kind == "adam" else None
    w_view = wrap_torch_tensor(w[:, :dim])  # sizes [ROWS, dim], strides [stride, 1]: what map_local_tensor hands the reference
    state_t = wrap_torch_tensor(state) if state is not None else None
    b12_t = wrap_torch_tensor(b12) if b12 is not None else None
    env = get_wholegraph_env_fns()
    for idx, g in steps:
        it, gt = torch.from_numpy(idx).cuda(), torch.from_numpy(g).cuda()
        wi, wg = wrap_torch_tensor(it), wrap_torch_tensor(gt)
        deduped = ctypes.c_int64(-1)
        rc = fn(OPT_ID[kind], wi.get_c_handle(), wg.get_c_handle(), w_view.get_c_handle(),
                state_t.get_c_handle() if state_t else None, b12_t.get_c_handle() if b12_t else None, 0,
                p["weight_decay"], p["epsilon"], p["beta1"], p["beta2"], 1 if p["adam_w"] > 0.5 else 0, p["alpha"], LR,
                env, get_stream(), ctypes.byref(deduped))
        assert rc == 0, "reference optimizer hook failed with code %d" % rc
        assert deduped.value == len(np.unique(idx)), (deduped.value, len(np.unique(idx)))
    torch.cuda.synchronize()
    res = {"w": w[:, :dim].cpu().numpy().copy()}
    if kind == "adam":
        res["m"] = state[:, :dim].cpu().numpy().copy()
        res["v"] = state[:, stride:stride + dim].cpu().numpy().copy()
        res["beta12t"] = b12.cpu().numpy().copy()
    elif kind == "adagrad":
        res["state_sum"] = state[:, :dim].cpu().numpy().copy()
    elif kind == "rmsprop":
        res["v"] = state[:, :dim].cpu().numpy().copy()
    return res


def run_all(out_path):
    """WG_OPT_WORKER_MODE: "api" (public C ABI; default for this repo's library), "hook" (the reference's dedup + optimizer
    kernels through oracle/ref_optimizer_hook.cpp; default when WHOLEGRAPH_B200_LIB is set)."""
    ref_lib = bool(os.environ.get("WHOLEGRAPH_B200_LIB"))
    mode = os.environ.get("WG_OPT_WORKER_MODE", "hook" if ref_lib else "api")
    out = {}
    for ci in range(len(CASES)):
        res = run_reference(ci) if mode == "hook" else run_api(ci, with_states=not ref_lib)
        for k, v in res.items():
            out["case%d_%s" % (ci, k)] = v
    np.savez_compressed(out_path, **out)


if __name__ == "__main__":
    run_all(sys.argv[1])
    print("optimizer worker done:", os.environ.get("WHOLEGRAPH_B200_LIB", "libwholegraph.so (this repo)"))
