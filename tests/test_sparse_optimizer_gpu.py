"""GPU tests of the fused duplicate-merge + sparse-optimizer kernel (wholememory_embedding_gather_gradient_apply at
world_size 1) against the oracle's restatement of the reference's dedup + optimizer kernels
(exchange_embeddings_nccl_func.cu:76-206, embedding_optimizer_func.cu:178-851).

Tolerance: fp32 results within rtol=1e-5 / atol=1e-5 of the oracle (the oracle is compiled C; FMA contraction differs
from nvcc's, so bit equality is not expected -- bit equality with the reference's own GPU kernels is what
test_ref_parity_gpu.py checks where the reference build covers the op).  Rows that receive no valid gradient must be
bit-identical to their initial value."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _run_case(kind, params, dim, idx_dtype, n, rows=3000, steps=3, poison=True, seed=0, mem_type="distributed"):
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    comm = G.single_comm()
    rng = np.random.default_rng(seed)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, G.MT[mem_type], wmb.MlDevice, wmb.create_non_cache_policy())
    opt = wmb.create_optimizer({"sgd": wmb.OptSgd, "adam": wmb.OptLazyAdam, "adagrad": wmb.OptAdaGrad, "rmsprop": wmb.OptRmsProp}[kind], params)
    opt.add_embedding(emb)
    wt = emb.get_embedding_tensor()
    local, off = wt.get_local_tensor(wmb.MlDevice, 0)
    assert off == 0 and tuple(local.shape) == (rows, dim)
    w = rng.standard_normal((rows, dim)).astype(np.float32)
    local.copy_(torch.from_numpy(w))
    w_init = w.copy()
    m = np.zeros_like(w)
    v = np.zeros_like(w)
    b12 = np.ones((rows, 2), np.float32)
    touched = np.zeros(rows, bool)
    env = get_wholegraph_env_fns()
    for step in range(steps):
        idx = (rng.zipf(1.2, size=n) % rows).astype(np.int64) if n else np.zeros(0, np.int64)  # heavy duplication
        if poison and n >= 16:
            idx[::7] = -1 - rng.integers(0, 1 << 20, size=idx[::7].shape[0])  # negative ids: ignored
            idx[3::11] = rows + rng.integers(0, 1 << 20, size=idx[3::11].shape[0])  # past the end: ignored
        g = rng.standard_normal((n, dim)).astype(np.float32)
        idx_t = torch.from_numpy(idx.astype(idx_dtype)).cuda() if n else torch.empty(0, dtype=torch.int64 if idx_dtype == np.int64 else torch.int32, device="cuda")
        g_t = torch.from_numpy(g).cuda() if n else torch.empty(0, dim, device="cuda")
        wmb.EmbeddingGatherGradientApply(emb, wrap_torch_tensor(idx_t), wrap_torch_tensor(g_t), False, 0.05, env, get_stream())
        valid = (idx >= 0) & (idx < rows)
        touched[idx[valid]] = True
        urows, ug = O.dedup_gradients(idx[valid], g[valid])
        kw = dict(weight_decay=params.get("weight_decay", 0.0), epsilon=params.get("epsilon", 1e-8))
        if kind == "adam":
            O.optimizer_step("adam", w, urows, ug, 0.05, state=(m, v), b12=b12, adam_w=params.get("adam_w", 0) > 0.5,
                             beta1=params.get("beta1", 0.9), beta2=params.get("beta2", 0.999), **kw)
        elif kind == "sgd":
            O.optimizer_step("sgd", w, urows, ug, 0.05, weight_decay=kw["weight_decay"])
        elif kind == "adagrad":
            O.optimizer_step("adagrad", w, urows, ug, 0.05, state=m, **kw)
        else:
            O.optimizer_step("rmsprop", w, urows, ug, 0.05, state=m, alpha=params.get("alpha", 0.99), **kw)
    torch.cuda.synchronize()
    got = local.cpu().numpy()
    try:
        assert np.allclose(got, w, rtol=1e-5, atol=1e-5), f"{kind} dim={dim}: max abs err {np.abs(got - w).max()}"
        assert got[~touched].tobytes() == w_init[~touched].tobytes(), "rows without a valid gradient were modified"
        if kind == "adam":
            bt = emb.get_optimizer_state("beta12t").get_local_tensor(wmb.MlDevice, 0)[0].cpu().numpy()
            assert np.allclose(bt, b12, rtol=1e-6, atol=0), "per-row beta^t state differs"
            mt = emb.get_optimizer_state("m").get_local_tensor(wmb.MlDevice, 0)[0].cpu().numpy()
            assert np.allclose(mt, m, rtol=1e-5, atol=1e-6)
    finally:
        emb.destroy_embedding()
        opt.destroy_optimizer()


@pytest.mark.parametrize("kind,params", [("sgd", {"weight_decay": 0.02}), ("adam", {}), ("adam", {"adam_w": 1.0, "weight_decay": 0.01, "beta1": 0.8}),
                                         ("adagrad", {"epsilon": 1e-6}), ("rmsprop", {"alpha": 0.9, "weight_decay": 0.001})])
@pytest.mark.parametrize("dim", [1, 4, 127, 128, 132, 392, 513, 1024])
def test_optimizers_all_dims_with_duplicates_and_invalid_ids(kind, params, dim):
    _run_case(kind, params, dim, np.int64, n=2500, seed=dim)


@pytest.mark.parametrize("idx_dtype", [np.int32, np.int64])
@pytest.mark.parametrize("mem_type", ["continuous", "chunked", "distributed"])
def test_lazy_adam_index_types_and_memory_types(idx_dtype, mem_type):
    _run_case("adam", {}, 64, idx_dtype, n=4000, seed=9, mem_type=mem_type)


@pytest.mark.parametrize("n", [0, 1, 31])
def test_lazy_adam_tiny_and_empty_batches(n):
    _run_case("adam", {}, 48, np.int64, n=n, seed=n, poison=False)


def test_all_gradients_on_one_row_sum_in_arrival_order():
    """2000 gradients for one row: the run walk adds them left to right like the reference's dedup kernel."""
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    comm = G.single_comm()
    rows, dim, n = 64, 256, 2000
    rng = np.random.default_rng(4)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, wmb.MtContinuous, wmb.MlDevice, wmb.create_non_cache_policy())
    opt = wmb.create_optimizer(wmb.OptSgd, {})
    opt.add_embedding(emb)
    local, _ = emb.get_embedding_tensor().get_local_tensor(wmb.MlDevice, 0)
    local.zero_()
    g = (rng.standard_normal((n, dim)) * np.float32(10.0) ** rng.integers(-3, 4, size=(n, 1))).astype(np.float32)
    idx = np.full(n, 17, np.int64)
    wmb.EmbeddingGatherGradientApply(emb, wrap_torch_tensor(torch.from_numpy(idx).cuda()), wrap_torch_tensor(torch.from_numpy(g).cuda()),
                                     False, 1.0, get_wholegraph_env_fns(), get_stream())
    torch.cuda.synchronize()
    acc = np.zeros(dim, np.float32)
    for i in range(n):  # sequential fp32 sum, arrival order
        acc = acc + g[i]
    got = local.cpu().numpy()
    # w = 0 - lr * sum  (weight_decay 0): bit-exact, additions only
    assert got[17].tobytes() == (np.float32(0) - acc).astype(np.float32).tobytes()
    assert not got[np.arange(rows) != 17].any()
    emb.destroy_embedding()
    opt.destroy_optimizer()


@pytest.mark.parametrize("kind,params", [("sgd", {"weight_decay": 0.02}), ("adam", {"weight_decay": 0.01, "beta1": 0.85}), ("adagrad", {}), ("rmsprop", {"alpha": 0.9})])
@pytest.mark.parametrize("dim", [4, 127, 512, 1028])
def test_hot_rows_take_the_long_run_kernel(kind, params, dim):
    """Rows with more than 64 gradients in one step are merged by long_run_update_kernel (a CTA per run, deep prefetch, same
    arrival-order sum).  Run lengths on both sides of the threshold and of its multiples (64, 65, 128, 129, 1000, 5000), heads
    at every alignment relative to the 64-position grid, mixed with short runs and singletons, two steps; results within
    1e-5 of the oracle like every other optimizer case, and SGD's merged sum bit-exact (additions only)."""
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    comm = G.single_comm()
    rows = 4000
    rng = np.random.default_rng(dim)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, wmb.MtChunked, wmb.MlDevice, wmb.create_non_cache_policy())
    opt = wmb.create_optimizer({"sgd": wmb.OptSgd, "adam": wmb.OptLazyAdam, "adagrad": wmb.OptAdaGrad, "rmsprop": wmb.OptRmsProp}[kind], params)
    opt.add_embedding(emb)
    local, _ = emb.get_embedding_tensor().get_local_tensor(wmb.MlDevice, 0)
    w = rng.standard_normal((rows, dim)).astype(np.float32)
    local.copy_(torch.from_numpy(w))
    m, v, b12 = np.zeros_like(w), np.zeros_like(w), np.ones((rows, 2), np.float32)
    env = get_wholegraph_env_fns()
    try:
        for step in range(2):
            runs = {7: 64, 8: 65, 100: 128, 101: 129, 900: 1000, 2500: 5000, 3999: 63, 0: 66}   # row id -> number of gradients
            ids = np.concatenate([np.full(c, r, np.int64) for r, c in runs.items()] + [rng.integers(0, rows, size=37 + step).astype(np.int64)])
            rng.shuffle(ids)   # arrival order is arbitrary; the sort is stable
            # magnitudes in [0.01, 1]: sums of thousands of rows stay small enough that cancellation in w - lr * g does not turn
            # the one-ulp FMA differences between nvcc and the C oracle into relative errors above the 1e-5 bar
            g = (rng.standard_normal((ids.size, dim)) * np.float32(10.0) ** rng.integers(-2, 1, size=(ids.size, 1))).astype(np.float32)
            idx_t, g_t = torch.from_numpy(ids).cuda(), torch.from_numpy(g).cuda()
            wmb.EmbeddingGatherGradientApply(emb, wrap_torch_tensor(idx_t), wrap_torch_tensor(g_t), False, 0.05, env, get_stream())
            urows, ug = O.dedup_gradients(ids, g)
            kw = dict(weight_decay=params.get("weight_decay", 0.0), epsilon=params.get("epsilon", 1e-8))
            if kind == "adam":
                O.optimizer_step("adam", w, urows, ug, 0.05, state=(m, v), b12=b12, beta1=params.get("beta1", 0.9), **kw)
            elif kind == "sgd":
                O.optimizer_step("sgd", w, urows, ug, 0.05, weight_decay=kw["weight_decay"])
            elif kind == "adagrad":
                O.optimizer_step("adagrad", w, urows, ug, 0.05, state=m, **kw)
            else:
                O.optimizer_step("rmsprop", w, urows, ug, 0.05, state=m, alpha=params.get("alpha", 0.99), **kw)
        torch.cuda.synchronize()
        got = local.cpu().numpy()
        assert np.allclose(got, w, rtol=1e-5, atol=1e-5), f"{kind} dim={dim}: max abs err {np.abs(got - w).max()} in rows {np.unique(np.argwhere(~np.isclose(got, w, rtol=1e-5, atol=1e-5))[:, 0])[:10]}"
        if kind == "adam":
            bt = emb.get_optimizer_state("beta12t").get_local_tensor(wmb.MlDevice, 0)[0].cpu().numpy()
            assert np.allclose(bt, b12, rtol=1e-6, atol=0), "per-row beta^t state differs"
    finally:
        emb.destroy_embedding()
        opt.destroy_optimizer()


def test_five_thousand_gradients_on_one_row_sum_bit_exactly():
    """The long-run kernel keeps the sequential left-to-right sum: SGD with lr 1 and w 0 returns minus that sum, bit for bit."""
    import gpu_utils as G
    import wholegraph_b200.binding as wmb
    from wholegraph_b200.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
    comm = G.single_comm()
    rows, dim, n = 64, 516, 5000   # 516 floats: a second column pass for 4 of the 128 threads
    rng = np.random.default_rng(44)
    d = wmb.PyWholeMemoryTensorDescription()
    d.set_dtype(wmb.DtFloat)
    d.set_shape((rows, dim))
    d.set_stride((dim, 1))
    emb = wmb.create_embedding(d, comm, wmb.MtContinuous, wmb.MlDevice, wmb.create_non_cache_policy())
    opt = wmb.create_optimizer(wmb.OptSgd, {})
    opt.add_embedding(emb)
    local, _ = emb.get_embedding_tensor().get_local_tensor(wmb.MlDevice, 0)
    local.zero_()
    g = (rng.standard_normal((n, dim)) * np.float32(10.0) ** rng.integers(-3, 4, size=(n, 1))).astype(np.float32)
    idx = np.full(n, 17, np.int64)
    idx_t, g_t = torch.from_numpy(idx).cuda(), torch.from_numpy(g).cuda()
    wmb.EmbeddingGatherGradientApply(emb, wrap_torch_tensor(idx_t), wrap_torch_tensor(g_t), False, 1.0, get_wholegraph_env_fns(), get_stream())
    torch.cuda.synchronize()
    acc = np.zeros(dim, np.float32)
    for i in range(n):
        acc = acc + g[i]
    got = local.cpu().numpy()
    assert got[17].tobytes() == (np.float32(0) - acc).astype(np.float32).tobytes()
    assert not got[np.arange(rows) != 17].any()
    emb.destroy_embedding()
    opt.destroy_optimizer()
