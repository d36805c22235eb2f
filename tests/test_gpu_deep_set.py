"""GPU parity of the Deep-Set conditioner Phi (examples/deep_set.ipynb:138-160) and of the joint
DeepSetFlow train step (deep_set.ipynb:313-352) against the CPU oracle / float64 autograd."""
import numpy as np
import pytest
import torch

from oracle import deep_set_oracle as dso
from oracle import zenflow_oracle as zo
from tests.helpers import to64

pytestmark = pytest.mark.gpu


def _data(seed=3, n_sets=37, max_size=40, pad=64):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, max_size, n_sets)
    n = int(sizes.sum())
    x = np.concatenate([rng.normal(size=(n, 2)), np.zeros((pad, 2))]).astype(np.float32)   # padded like preprocess()
    set_idx, row_idx = dso.coo_from_sizes(sizes)
    return x, sizes, set_idx, row_idx


def _sum_matrix(sizes):
    from zenflow_b200.deep_set import SumMatrix

    return SumMatrix.from_sizes(sizes)


def test_phi_eval_matches_oracle():
    from zenflow_b200.deep_set import Phi

    x, sizes, set_idx, row_idx = _data()
    v = dso.init_phi(seed=5)
    c = Phi().apply(v, x, _sum_matrix(sizes))
    c64, _ = dso.phi_forward(to64(v), x.astype(np.float64), set_idx, row_idx, len(sizes))
    c32, _ = dso.phi_forward(v, x, set_idx, row_idx, len(sizes))
    assert c.shape == (len(sizes), 8)
    err, ref = np.abs(c - c64).max(), np.abs(c32 - c64).max()
    assert err <= 4 * ref + 1e-5 * np.abs(c64).max(), (err, ref)


def test_phi_train_forward_explicit_mask_and_running_stats():
    from zenflow_b200.deep_set import Phi

    x, sizes, set_idx, row_idx = _data(seed=4)
    v = dso.init_phi(seed=6)
    rng = np.random.default_rng(0)
    mult = (rng.uniform(size=(x.shape[0], 8)) >= 0.3).astype(np.float32) / 0.7
    c, upd = Phi().apply(v, x, _sum_matrix(sizes), train=True, dropout_mask=mult)
    c64, st64 = dso.phi_forward(to64(v), x.astype(np.float64), set_idx, row_idx, len(sizes), train=True, dropout_mult=mult)
    assert np.abs(c - c64).max() <= 2e-5 * np.abs(c64).max() + 1e-5
    got = upd["batch_stats"]["BatchNorm_0"]
    np.testing.assert_allclose(got["mean"], st64["BatchNorm_0"]["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got["var"], st64["BatchNorm_0"]["var"], rtol=1e-5, atol=1e-6)


def test_phi_philox_dropout_keeps_the_right_fraction_and_is_reproducible():
    from zenflow_b200.deep_set import Phi

    x, sizes, _, _ = _data(seed=8, n_sets=200, max_size=60)
    v = dso.init_phi(seed=1)
    phi, sm = Phi(), _sum_matrix(sizes)
    a, _ = phi.apply(v, x, sm, train=True, seed=11)
    b, _ = phi.apply(v, x, sm, train=True, seed=11)
    d, _ = phi.apply(v, x, sm, train=True, seed=12)
    # same seed -> same keep-mask (the pooling adds with fp32 atomics, so only the summation order may differ)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5)
    assert np.abs(a - d).max() > 1e-2
    # E[dropout(h)] = h: the pooled train output (with batch-stat BN) averages to the rate-0 train output
    nodrop, _ = Phi(rate=0.0).apply(v, x, sm, train=True)
    many = np.mean([phi.apply(v, x, sm, train=True, seed=s)[0] for s in range(40)], axis=0)
    assert np.abs(many - nodrop).mean() < 0.15 * np.abs(nodrop).mean()


def _torch_phi(v, x, set_idx, row_idx, S, mult):
    dt = torch.float64
    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=dt, requires_grad=True)
    p = {"scale": t(v["params"]["BatchNorm_0"]["scale"]), "bias": t(v["params"]["BatchNorm_0"]["bias"])}
    dense = [(t(d["kernel"]), t(d["bias"])) for _, d in sorted(v["params"]["NNBlock_0"].items())]
    xt = torch.tensor(x, dtype=dt)
    mean, mean2 = xt.mean(0), (xt * xt).mean(0)
    var = torch.clamp(mean2 - mean * mean, min=0)
    h = (xt - mean) * (1.0 / torch.sqrt(var + 1e-5) * p["scale"]) + p["bias"]
    for j, (k, b) in enumerate(dense):
        h = h @ k + b
        if j < len(dense) - 1:
            h = h * torch.sigmoid(h)
    h = h * torch.tensor(mult, dtype=dt)
    c = torch.zeros(S, h.shape[1], dtype=dt).index_add(0, torch.tensor(set_idx, dtype=torch.long), h[torch.tensor(row_idx, dtype=torch.long)])
    return c, p, dense


def test_phi_backward_matches_float64_autograd():
    from zenflow_b200.deep_set import Phi, PhiEngine

    x, sizes, set_idx, row_idx = _data(seed=9, n_sets=50, max_size=30)
    v = dso.init_phi(seed=2)
    rng = np.random.default_rng(1)
    mult = (rng.uniform(size=(x.shape[0], 8)) >= 0.3).astype(np.float32) / 0.7
    gc = rng.normal(size=(len(sizes), 8)).astype(np.float32)
    c_t, p_t, dense_t = _torch_phi(v, x, set_idx, row_idx, len(sizes), mult)
    (c_t * torch.tensor(gc, dtype=torch.float64)).sum().backward()
    eng = PhiEngine(Phi(), v, 2)
    c = eng.forward(x, _sum_matrix(sizes), train=True, dropout_mask=mult)
    assert np.abs(c.cpu().numpy() - c_t.detach().numpy()).max() <= 2e-5 * np.abs(c_t.detach().numpy()).max() + 1e-5
    eng.backward(torch.from_numpy(gc).cuda())
    g = eng.gradients()
    checks = [("scale", g["BatchNorm_0"]["scale"], p_t["scale"].grad), ("bias", g["BatchNorm_0"]["bias"], p_t["bias"].grad)]
    for j, (k, b) in enumerate(dense_t):
        checks += [(f"Dense_{j}/kernel", g["NNBlock_0"][f"Dense_{j}"]["kernel"], k.grad),
                   (f"Dense_{j}/bias", g["NNBlock_0"][f"Dense_{j}"]["bias"], b.grad)]
    for name, got, ref in checks:
        ref = ref.numpy()
        e = np.abs(got.cpu().numpy() - ref).max() / (np.abs(ref).max() + 1e-12)
        assert e <= 1e-4, f"{name}: rel err {e:.2e}"


def test_deep_set_flow_joint_step_learns():
    """deep_set.ipynb:313-352 in miniature: Phi's conditions reach the flow, d loss / d c reaches Phi, the loss falls."""
    from zenflow_b200 import Flow
    from zenflow_b200.bijectors import rolling_spline_coupling
    from zenflow_b200.deep_set import DeepSetFlowTrainer, Phi, SumMatrix
    from zenflow_b200.distributions import Beta

    rng = np.random.default_rng(1)
    n_sets = 300
    sizes = (rng.exponential(size=n_sets) * 40).astype(int) + 1
    X = np.concatenate([rng.normal(size=(s, 2)) for s in sizes]).astype(np.float32)
    y = rng.normal(np.sqrt(sizes), 1, size=(2, n_sets)).T.astype(np.float32)
    phi = Phi()
    flow = Flow(rolling_spline_coupling(2, layers=(128,) * 6), latent=Beta())
    pv = phi.init(0, X)
    fv = flow.init(0, y[:1], np.zeros((1, 8), np.float32))
    tr = DeepSetFlowTrainer(phi, pv, flow, fv, 2, 2)
    sm = SumMatrix.from_sizes(sizes)
    losses = []
    for epoch in range(60):
        lp_sum = tr.step(X, sm, y, seed=epoch)
        losses.append(-float(lp_sum.item()) / n_sets)
    assert np.isfinite(losses).all()
    assert min(losses[-10:]) < losses[0] - 0.3, losses[::10]
    lp = tr.log_prob(X, sm, y)
    assert lp.shape == (n_sets,) and bool(torch.isfinite(lp).all())


def test_sum_matrix_indices_are_validated():
    """The kernels index c and h with the COO lists: out-of-range entries must raise on the host, not corrupt memory."""
    from zenflow_b200.deep_set import Phi, SumMatrix

    sm = SumMatrix(np.array([0, 0, 1, 2]), np.array([0, 1, 2, 3]), 3)   # numpy / int64 input is accepted
    assert sm.set_idx.dtype == torch.int32 and sm.set_idx.is_cuda and sm.nnz == 4
    with pytest.raises(ValueError, match="set indices must be in"):
        SumMatrix(np.array([0, 3]), np.array([0, 1]), 3)
    with pytest.raises(ValueError, match="negative row index"):
        SumMatrix(np.array([0, 1]), np.array([0, -1]), 3)
    with pytest.raises(ValueError, match="row indices"):
        SumMatrix(np.array([0, 1]), np.array([0]), 3)
    phi = Phi()
    x = np.zeros((3, 2), np.float32)   # 3 rows, the matrix refers to row 3
    with pytest.raises(ValueError, match="out of bounds for 3 rows"):
        phi.apply(phi.init(0, x), x, sm)
