"""GPU parity of the train-mode path: batch statistics, the train step's loss and gradients
(against float64 autograd of the oracle = the role jax.grad plays in train.py:82), the NAdamW
update, and the ``train`` loop."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as to
from oracle import zenflow_oracle as zo
from tests.helpers import product_chain, to64

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-4  # north_star: gradients within rel 1e-4 (of the leaf's largest entry)


def _flow_vars(v):
    return {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}


def test_shift_bounds_train_golden():
    """tests/test_bijectors.py:35-58 through the CUDA path: running stats, output, inverse."""
    from zenflow_b200 import bijectors as bi

    x = np.array([[1, 5], [3, 4], [6, 2]])
    sb = bi.ShiftBounds(margin=0.01)
    variables = sb.init(0, x, None)
    (y, log_det), updates = sb.apply(variables, x, None, train=True, mutable=["batch_stats"])
    bs = updates["batch_stats"]
    np.testing.assert_allclose(bs["xmin_0"], 0.975, rtol=1e-6)
    np.testing.assert_allclose(bs["xmax_0"], 6.025, rtol=1e-6)
    np.testing.assert_allclose(bs["xmin_1"], 1.985, rtol=1e-6)
    np.testing.assert_allclose(bs["xmax_1"], 5.015, rtol=1e-6)
    assert np.isposinf(variables["batch_stats"]["xmin_0"]).all()  # input tree untouched
    y_ref = np.column_stack([(x[:, i] - bs[f"xmin_{i}"]) / (bs[f"xmax_{i}"] - bs[f"xmin_{i}"]) for i in range(2)])
    np.testing.assert_allclose(y, y_ref, atol=5e-6)
    x2 = sb.apply(updates, y, None, method="inverse")
    np.testing.assert_allclose(x2, x, atol=6e-6)
    with pytest.raises(ValueError):  # immutable collection
        sb.apply(variables, x, None, train=True)


def test_chain_shiftbounds_roll_train_golden():
    """tests/test_bijectors.py:191-206."""
    from zenflow_b200 import bijectors as bi

    x = np.array([[2.5, 2, 3], [1, 3.5, 4.5], [4, 5, 6]], dtype=np.float32)
    chain = bi.Chain([bi.ShiftBounds(margin=0.0), bi.Roll()])
    variables = chain.init(0, x, None)
    (y, log_det), updates = chain.apply(variables, x, None, train=True, mutable=["batch_stats"])
    np.testing.assert_allclose(y, [[0.0, 0.5, 0.0], [0.5, 0.0, 0.5], [1.0, 1.0, 1.0]], atol=1e-6)
    ld_ref = chain[0].apply({"batch_stats": updates["batch_stats"]["bijectors_0"]}, x, None)[1]
    np.testing.assert_allclose(log_det, ld_ref, atol=5e-6)
    np.testing.assert_allclose(chain.apply(updates, y, None, method="inverse"), x, atol=1e-6)


@pytest.mark.parametrize("D,C,K,layers,M", [(2, 1, 16, (128, 128), 1000), (5, 0, 6, (32, 24), 777)])
def test_flow_train_mode_forward(D, C, K, layers, M):
    """Flow.apply(train=True, mutable=['batch_stats']) (train.py:66-72): log-prob with batch
    statistics and the updated running statistics."""
    from zenflow_b200 import Flow

    rng = np.random.default_rng(D)
    ops = zo.make_chain(D, K, layers)
    x = rng.normal(0.2, 1.1, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
    v = zo.init_variables(ops, D, C, 3, weight_scale=1.5, randomize_bn=True)
    flow = Flow(product_chain(ops))
    flow.latent._latch_dim(D)
    lp, upd = flow.apply(_flow_vars(v), x, c, train=True, mutable=["batch_stats"])
    lp64, st64 = zo.flow_log_prob(ops, to64(v), x.astype(np.float64), None if c is None else c.astype(np.float64), train=True)
    lp32, _ = zo.flow_log_prob(ops, v, x, c, train=True)
    err, ref = np.abs(lp - lp64), np.abs(lp32 - lp64)
    assert err.max() <= 2 * ref.max() + 2e-5, (err.max(), ref.max())
    got = upd["batch_stats"]["bijector"]
    for name, st in st64.items():
        if "BatchNorm_0" in st:
            np.testing.assert_allclose(got[name]["BatchNorm_0"]["mean"], st["BatchNorm_0"]["mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(got[name]["BatchNorm_0"]["var"], st["BatchNorm_0"]["var"], rtol=1e-5, atol=1e-6)
        else:
            for k2, val in st.items():
                np.testing.assert_allclose(got[name][k2], val, rtol=1e-6)


GRAD_CASES = [
    # D, C, K, layers, n_couplings, roll, M, weight scale (1.0 = the FLAX default initialiser's variance)
    (4, 2, 8, (16, 16), None, 1, 300, 1.5),
    (2, 1, 16, (128, 128), None, 1, 700, 1.5),
    (16, 4, 32, (128, 128), 3, 2, 260, 1.0),
    (3, 0, 5, (40,), None, 1, 515, 1.5),
    (24, 8, 16, (128, 128), 2, 3, 300, 1.0),   # fused VJP kernel: 20 conditioner inputs (FFMA first Dense), K = 16 rows
    (6, 1, 32, (128, 128), 2, 1, 400, 1.0),    # fused VJP kernel: 3 transformed dims, tensor-core first Dense of 4 inputs
]


@pytest.mark.parametrize("case", GRAD_CASES, ids=[f"D{c[0]}C{c[1]}K{c[2]}" for c in GRAD_CASES])
def test_train_step_gradients_match_autograd(case):
    from zenflow_b200 import Flow
    from zenflow_b200._train import TrainEngine

    D, C, K, layers, ncoup, roll, M, wscale = case
    rng = np.random.default_rng(M)
    ops = zo.make_chain(D, K, layers, n_couplings=ncoup, roll_shift=roll)
    x = rng.normal(0.3, 1.0, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
    v = zo.init_variables(ops, D, C, 2, weight_scale=wscale, randomize_bn=True)
    loss64, g64, st64, gc64, lp64 = to.loss_and_grads(ops, to64(v), x.astype(np.float64),
                                                      None if c is None else c.astype(np.float64))
    flow = Flow(product_chain(ops))
    flow.latent._latch_dim(D)
    eng = TrainEngine(flow, _flow_vars(v), D, C, micro_batch=128)  # several micro-batches incl. a ragged one
    lp_sum, gc_dev = eng.step(x, c, update=False, want_gc=True)
    loss = -float(lp_sum.item()) / M
    assert abs(loss - loss64) <= 1e-5 * abs(loss64) + 1e-5
    grads = eng.gradients()["bijector"]
    worst = 0.0
    for name, layers_ in g64.items():
        for lname, leaves in layers_.items():
            for leaf, ref in leaves.items():
                got = grads[name][lname][leaf].cpu().numpy()
                scale = np.abs(ref).max() + 1e-12
                e = np.abs(got - ref).max() / scale
                worst = max(worst, e)
                assert e <= GRAD_RTOL, f"{name}/{lname}/{leaf}: rel err {e:.2e} (scale {scale:.2e})"
    print(f"\nworst gradient rel err {worst:.2e}; loss {loss:.6f} vs {loss64:.6f}")
    if C:
        gc = gc_dev.cpu().numpy()
        assert np.abs(gc - gc64).max() <= GRAD_RTOL * np.abs(gc64).max()
    # running statistics after the step (train.py:83)
    vs = eng.variables(as_numpy=True)["batch_stats"]["bijector"]
    for name, st in st64.items():
        if "BatchNorm_0" in st:
            np.testing.assert_allclose(vs[name]["BatchNorm_0"]["var"], st["BatchNorm_0"]["var"], rtol=2e-5, atol=1e-6)


def test_nadamw_update_matches_optax_formula():
    from zenflow_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = 10_007
    p = rng.normal(size=n).astype(np.float32)
    mu = np.zeros(n); nu = np.zeros(n); pr = p.astype(np.float64)
    pt = torch.from_numpy(p.copy()).cuda()
    mut, nut = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for nesterov in (True, False):
        count = 0
        for it in range(4):
            g = rng.normal(scale=0.3, size=n).astype(np.float32)
            pr, mu, nu, _ = to.nadamw_update(pr, g.astype(np.float64), mu, nu, count, nesterov=nesterov)
            gt = torch.from_numpy(g).cuda()
            _lib.check(lib.zf_nadamw_update(torch.cuda.current_stream().cuda_stream, n, pt.data_ptr(), gt.data_ptr(),
                                            mut.data_ptr(), nut.data_ptr(), count, 1e-3, 0.9, 0.999, 1e-8, 1e-4,
                                            int(nesterov)))
            count += 1
            np.testing.assert_allclose(pt.cpu().numpy(), pr, rtol=2e-6, atol=2e-7)


def test_train_loop_two_moons():
    """train() end to end (reference: tests/test_train.py, examples/two_moons.ipynb): the loss
    falls well below the untrained value and the bookkeeping has the reference's shape."""
    from zenflow_b200 import Flow, train
    from zenflow_b200.bijectors import rolling_spline_coupling

    rng = np.random.default_rng(1)
    n = 4000
    lab = rng.integers(0, 2, n)
    t = rng.uniform(0, np.pi, n)
    X = np.column_stack([np.where(lab == 0, np.cos(t), 1 - np.cos(t)), np.where(lab == 0, np.sin(t), 0.5 - np.sin(t))])
    X = (X + 0.1 * rng.standard_normal(X.shape)).astype(np.float32)
    flow = Flow(rolling_spline_coupling(2))
    best, best_epoch, ltrain, ltest = train(flow, X[:3000], X[3000:], epochs=12, batch_size=500, patience=4, progress=False)
    print("\ntwo_moons test loss per epoch:", [round(v, 3) for v in ltest])
    assert len(ltrain) == len(ltest) == 12 and np.isfinite(ltrain).all()
    assert min(ltest) < ltest[0] - 0.2
    assert 0 <= best_epoch < 12 and set(best) == {"params", "batch_stats"}
    xs = flow.apply(best, 1000, method="sample")
    assert xs.shape == (1000, 2) and bool(torch.isfinite(xs).all())
    # conditional variant + ragged last minibatch
    C = lab.astype(np.float32)
    flow = Flow(rolling_spline_coupling(2))
    best, best_epoch, ltrain, ltest = train(flow, X[:3001], X[3001:], C[:3001], C[3001:], epochs=4, batch_size=512,
                                            patience=2, progress=False)
    assert np.isfinite(ltrain).all() and ltest[-1] < ltest[0]


@pytest.mark.filterwarnings("error::RuntimeWarning")
def test_bad_input_distribution():
    """tests/test_train.py:7-15 (Pareto data), shortened from 1000 to 30 epochs."""
    from zenflow_b200 import Flow, train
    from zenflow_b200.bijectors import rolling_spline_coupling

    rng = np.random.default_rng(1)
    x = rng.pareto(5, size=1000)
    flow = Flow(rolling_spline_coupling(2))
    X = np.column_stack((x, x)).astype(np.float32)
    loss_train = train(flow, X, X, epochs=30, progress=False)[2]
    assert np.all(np.isfinite(loss_train))


def test_epoch_shuffle_is_a_permutation():
    """zf_permute_rows (X_train[perm], train.py:104-108): a permutation of the rows, keyed by the seed,
    identical row pairing for x and c."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for N, D in [(1, 3), (2, 1), (1000, 2), (4097, 16), (100_003, 5)]:
        x = torch.arange(N * D, dtype=torch.float32, device="cuda").reshape(N, D)
        c = torch.arange(N, dtype=torch.float32, device="cuda").reshape(N, 1)
        out, outc, out2 = torch.empty_like(x), torch.empty_like(c), torch.empty_like(x)
        _lib.check(lib.zf_permute_rows(st, x.data_ptr(), N, D, 42, out.data_ptr()))
        _lib.check(lib.zf_permute_rows(st, c.data_ptr(), N, 1, 42, outc.data_ptr()))
        _lib.check(lib.zf_permute_rows(st, x.data_ptr(), N, D, 43, out2.data_ptr()))
        rows = (out[:, 0] / D).long()
        assert torch.equal(torch.sort(rows).values, torch.arange(N, device="cuda"))     # a permutation
        assert torch.equal(out, x[rows]) and torch.equal(outc[:, 0].long(), rows)        # whole rows, same pairing
        if N > 100:
            assert not torch.equal(out, out2) and not torch.equal(rows, torch.arange(N, device="cuda"))
            # well mixed: neighbouring outputs come from far-apart inputs
            assert float((rows[1:] - rows[:-1]).abs().float().mean()) > N / 5
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    lp = torch.randn(100_001, device="cuda")
    _lib.check(lib.zf_neg_sum(st, lp.data_ptr(), lp.numel(), acc.data_ptr()))
    assert abs(acc.item() + lp.double().sum().item()) < 1e-6


def test_nadamw_device_counter_matches_host_counter():
    """zf_nadamw_update_dev (step counter and bias corrections on the device: what lets a train step be replayed as a
    CUDA graph) against zf_nadamw_update on the same gradients."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)
    n = 4099
    p0 = torch.from_numpy(rng.normal(size=n).astype(np.float32)).cuda()
    for nesterov in (1, 0):
        pa, pb = p0.clone(), p0.clone()
        ma, va, mb, vb = (torch.zeros(n, device="cuda") for _ in range(4))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        bias = torch.zeros(4, device="cuda")
        for it in range(7):
            g = torch.from_numpy(rng.normal(scale=0.3, size=n).astype(np.float32)).cuda()
            _lib.check(lib.zf_nadamw_update(st, n, pa.data_ptr(), g.data_ptr(), ma.data_ptr(), va.data_ptr(), it,
                                            1e-3, 0.9, 0.999, 1e-8, 1e-4, nesterov))
            _lib.check(lib.zf_nadamw_update_dev(st, n, pb.data_ptr(), g.data_ptr(), mb.data_ptr(), vb.data_ptr(),
                                                cnt.data_ptr(), bias.data_ptr(), 1e-3, 0.9, 0.999, 1e-8, 1e-4, nesterov))
            assert int(cnt.item()) == it + 1
            np.testing.assert_allclose(pb.cpu().numpy(), pa.cpu().numpy(), rtol=3e-7, atol=1e-9)


def test_graphed_small_batch_steps_match_eager_steps():
    """Small single-device steps replay a captured CUDA graph (TrainEngine._step_graphed): same losses, parameters,
    running statistics and step count as the eager sequence, over full and ragged minibatches with fresh data each
    step (a stale staging buffer or a frozen step counter would show)."""
    from zenflow_b200 import Flow, _lib
    from zenflow_b200._train import TrainEngine

    D, C, K, layers = 2, 1, 16, (128, 128)
    ops = zo.make_chain(D, K, layers)
    v = zo.init_variables(ops, D, C, 4, weight_scale=1.0, randomize_bn=True)
    rng = np.random.default_rng(11)
    sizes = [256, 256, 100, 256, 256, 100, 256, 256]
    batches = [(rng.normal(0.2, 1.0, (m, D)).astype(np.float32), rng.uniform(0, 1, (m, C)).astype(np.float32)) for m in sizes]
    engines = []
    for graphs in (False, True):
        flow = Flow(product_chain(ops))
        flow.latent._latch_dim(D)
        eng = TrainEngine(flow, _flow_vars(v), D, C)
        eng.use_graphs = graphs
        engines.append(eng)
    losses = [[], []]
    launches = []
    for x, c in batches:
        for i, eng in enumerate(engines):
            n0 = _lib.launch_count()
            losses[i].append(-float(eng.step(x, c).item()) / len(x))
            launches.append(_lib.launch_count() - n0)
    eager, graphed = engines
    assert any(e["graph"] is not None for e in graphed._graphs.values()) and not eager._graphs
    assert launches[0::2] == launches[1::2] and min(launches) > 10   # replays are counted like the eager launches
    assert graphed.count == eager.count == len(sizes) == int(graphed._count_dev.item()) == int(eager._count_dev.item())
    np.testing.assert_allclose(losses[1], losses[0], rtol=2e-5, atol=1e-6)
    assert losses[0][-1] < losses[0][0]
    pe, pg = eager.P.cpu().numpy(), graphed.P.cpu().numpy()
    diff = np.abs(pe - pg)
    # fp32 atomics order the gradient sums differently from run to run: Adam turns a near-zero gradient's noise into
    # a step of up to lr, so a handful of parameters may differ by that much; everything else agrees
    assert np.quantile(diff, 0.99) <= 1e-5 and diff.max() <= 1e-3 * len(sizes)
    ve, vg = eager.variables(as_numpy=True)["batch_stats"], graphed.variables(as_numpy=True)["batch_stats"]
    flat = lambda t: np.concatenate([np.ravel(x) for x in torch.utils._pytree.tree_leaves(t)])
    np.testing.assert_allclose(flat(vg), flat(ve), rtol=1e-4, atol=1e-5)
