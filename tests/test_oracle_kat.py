"""The reference's own known-answer tests, re-expressed against the CPU oracle.

Each test cites the reference test it restates (paths relative to /root/reference).
These pin the oracle; the CUDA parity tests then compare against the oracle.
"""
import math

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal
from scipy import stats as sps

from oracle import zenflow_oracle as zo


def test_rqs_identity():
    """tests/test_utils.py:7-13 — equal bins, unit slopes ⇒ identity, also out of range."""
    for dt in (np.float32, np.float64):
        x = np.linspace(-1, 2, 10).reshape(-1, 1).astype(dt)
        W = np.tile([0.25] * 4, len(x)).reshape(*x.shape, -1).astype(dt)
        D = np.tile([1.0] * 3, len(x)).reshape(*x.shape, -1).astype(dt)
        y, log_det = zo.rqs_forward(x, W, W.copy(), D)
        assert_allclose(y, x, atol=1e-5)
        assert log_det.shape == (10,)


def test_rqs_logdet_is_log_derivative_and_roundtrip():
    """tests/test_utils.py:16-50 — log_det vs numeric dy/dx (jacobi there, central
    differences in float64 here), y≈x for small params, inverse∘forward≈id."""
    rng = np.random.default_rng(1)
    x = np.linspace(-0.1, 1.1, 1000).reshape(1000, 1)
    scale, K = 0.1, 3
    dx, dy, slope = zo.normalize_spline_params(
        scale * rng.normal(size=K), scale * rng.normal(size=K), scale * rng.normal(size=K - 1)
    )
    n = x.size
    dx = np.tile(dx, n).reshape(*x.shape, -1)
    dy = np.tile(dy, n).reshape(*x.shape, -1)
    slope = np.tile(slope, n).reshape(*x.shape, -1)
    y, log_det = zo.rqs_forward(x, dx, dy, slope)
    h = 1e-6
    yp, _ = zo.rqs_forward(x + h, dx, dy, slope)
    ym, _ = zo.rqs_forward(x - h, dx, dy, slope)
    j = ((yp - ym) / (2 * h)).reshape(-1)
    assert_allclose(y, x, atol=0.1)
    assert_allclose(log_det, np.log(j), atol=0.01)
    x2 = zo.rqs_inverse(y, dx, dy, slope)
    assert_allclose(x2, x, atol=1e-4)
    # float32 arithmetic (the reference's) meets the same bounds
    y32, ld32 = zo.rqs_forward(x.astype(np.float32), dx.astype(np.float32),
                               dy.astype(np.float32), slope.astype(np.float32))
    assert y32.dtype == np.float32 and ld32.dtype == np.float32
    assert_allclose(ld32, np.log(j), atol=0.01)
    x2 = zo.rqs_inverse(y32, dx.astype(np.float32), dy.astype(np.float32), slope.astype(np.float32))
    assert_allclose(x2, x, atol=1e-4)


def test_index_golden():
    """tests/test_utils.py:53-68 — golden bin indices."""
    x = np.array([-2, -1, -0.5, -0.1, 0.0, 0.1, 0.5, 1.0, 1.5]).reshape(1, -1)
    xk = np.array([-1, 0, 1]).reshape(1, 3)
    ind, oob = zo.index(x, xk)
    assert_array_equal(ind[0, :, 0], [0, 0, 0, 0, 1, 1, 1, 2, 2])
    assert_array_equal(oob[0], [True, True, True, True, False, False, False, True, True])


def test_knots_golden():
    """tests/test_utils.py:71-74."""
    assert_allclose(zo.knots(np.array((0.25, 0.25, 0.25))), [0, 0.25, 0.5, 0.75])


@pytest.mark.parametrize("threshold", (0, 0.1))
def test_softmax_with_threshold_1(threshold):
    """tests/test_utils.py:77-83."""
    y = zo.softmax_with_threshold(np.array((-5.0, 1.0, 2.0)), threshold)
    assert_allclose(np.sum(y), 1)
    assert np.all(y >= threshold)


def test_softmax_with_threshold_2():
    """tests/test_utils.py:86-94."""
    y = zo.softmax_with_threshold(np.array([(-5.0, 1.0, 2.0), (-4.0, 2.0, 3.0)]), 0.1)
    assert_allclose(np.sum(y[0]), 1)
    assert_allclose(np.sum(y[1]), 1)
    assert np.all(y >= 0.1)


def test_shift_bounds_golden():
    """tests/test_bijectors.py:35-58 — running stats 0.975/6.025/1.985/5.015, output, inverse."""
    x = np.array([[1, 5], [3, 4], [6, 2]])
    st = {}
    zo.shift_bounds_forward(x, st, margin=0.01, train=False, initializing=True)  # init
    assert np.isinf(st["xmin_0"]).all() and np.isinf(st["xmax_1"]).all()
    y, log_det = zo.shift_bounds_forward(x, st, margin=0.01, train=True)
    assert_allclose(st["xmin_0"], 0.975)
    assert_allclose(st["xmax_0"], 6.025)
    assert_allclose(st["xmin_1"], 1.985)
    assert_allclose(st["xmax_1"], 5.015)
    y_ref = np.column_stack([(x[:, i] - st[f"xmin_{i}"]) / (st[f"xmax_{i}"] - st[f"xmin_{i}"])
                             for i in range(2)])
    assert_allclose(y, y_ref, atol=5e-6)
    x2 = zo.shift_bounds_inverse(y, st)
    assert_allclose(x2, x, atol=1e-6 * 6)  # fp32 ulp at 6 is 4.8e-7; the reference asserts 1e-6


def test_shift_bounds_bounded_variants():
    """tests/test_bijectors.py:61-92 — both/lower/upper bounds reference formulas."""
    rng = np.random.default_rng(0)
    x = np.column_stack([
        2 * rng.uniform(size=10) - 1,
        rng.exponential(size=10) * 10 + 10,
        1 - rng.exponential(size=10),
    ]).astype(np.float32)
    bounds = [(0, -1, 1), (1, 10, None), (2, None, 1)]
    st = {}
    y, _ = zo.shift_bounds_forward(x, st, margin=0.0, bounds=bounds, train=True)
    x2 = zo.shift_bounds_inverse(y, st, bounds=bounds)
    assert "xmin_0" not in st  # fully bounded column keeps no statistics
    t1 = np.log(x[:, 1] - 10)
    t2 = np.log(1 - x[:, 2])
    assert_allclose(y[:, 0], (x[:, 0] + 1) / 2, atol=1e-6)
    assert_allclose(y[:, 1], (t1 - t1.min()) / (t1.max() - t1.min()), atol=1e-6)
    assert_allclose(y[:, 2], (t2 - t2.min()) / (t2.max() - t2.min()), atol=1e-6)
    assert_allclose(x2, x, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("bound", [None, (0, -1, 1), (0, 10, None), (0, None, 1)])
def test_shift_bounds_logdet_numeric(bound):
    """tests/test_bijectors.py:95-165 — log-det vs numeric derivative (float64 here)."""
    rng = np.random.default_rng(3)
    if bound is None or bound == (0, -1, 1):
        x = 2 * rng.uniform(size=(20, 1)) - 1
    elif bound == (0, 10, None):
        x = rng.exponential(size=(20, 1)) * 10 + 10
    else:
        x = 1 - rng.exponential(size=(20, 1))
    bounds = [] if bound is None else [bound]
    st = {}
    y, _ = zo.shift_bounds_forward(x, st, bounds=bounds, train=True)
    keep = (y[:, 0] > 0.1) & (y[:, 0] < 0.9)
    x = x[keep]
    _, ld = zo.shift_bounds_forward(x, st, bounds=bounds)
    h = 1e-6
    yp, _ = zo.shift_bounds_forward(x + h, st, bounds=bounds)
    ym, _ = zo.shift_bounds_forward(x - h, st, bounds=bounds)
    assert_allclose(ld, np.log(np.abs((yp - ym)[:, 0] / (2 * h))), atol=1e-3)


def test_roll_golden():
    """tests/test_bijectors.py:168-176."""
    x = np.array([[1, 5], [3, 4], [6, 2]])
    z = zo.roll_forward(x)
    assert_array_equal(z, [[5, 1], [4, 3], [2, 6]])
    assert_array_equal(zo.roll_inverse(z), x)


def test_chain_of_rolls_golden():
    """tests/test_bijectors.py:179-188."""
    x = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32)
    ops = [{"kind": "roll", "shift": 1}, {"kind": "roll", "shift": 1}]
    z, ld, _ = zo.chain_forward(ops, {}, x, train=True)
    assert_array_equal(z, [[2, 3, 1], [5, 6, 4]])
    assert_array_equal(ld, np.zeros(2))
    assert_array_equal(zo.chain_inverse(ops, {}, z), x)


def test_chain_shiftbounds_roll_golden():
    """tests/test_bijectors.py:191-206 — exact outputs and the batch_stats sub-tree name."""
    x = np.array([[2.5, 2, 3], [1, 3.5, 4.5], [4, 5, 6]], dtype=np.float32)
    ops = [{"kind": "shift_bounds", "margin": 0.0, "bounds": ()}, {"kind": "roll", "shift": 1}]
    y, log_det, stats = zo.chain_forward(ops, {}, x, train=True)
    assert_allclose(y, [[0.0, 0.5, 0.0], [0.5, 0.0, 0.5], [1.0, 1.0, 1.0]])
    assert "bijectors_0" in stats and "xmin_0" in stats["bijectors_0"]
    _, ld_ref = zo.shift_bounds_forward(x, stats["bijectors_0"], margin=0.0)
    assert_allclose(log_det, ld_ref, atol=5e-6)
    x2 = zo.chain_inverse(ops, {"batch_stats": stats}, y)
    assert_allclose(x2, x)


def _train_then_eval(ops, x, c, seed=0):
    v = zo.init_variables(ops, x.shape[1], 0 if c is None else c.shape[1], seed)
    _, _, stats = zo.chain_forward(ops, v, x, c, train=True)
    v = {"params": v["params"], "batch_stats": stats}
    y, ld, _ = zo.chain_forward(ops, v, x, c, train=False)
    return v, y, ld


def test_chain_coupling_roundtrip():
    """tests/test_bijectors.py:209-226 (Chain_3) — inverse(forward(x)) ≈ x, rtol 1e-5
    in the reference with PRNGKey(0) weights; fp32 here with seed-0 weights, and the
    structural EPS mismatch between forward and inverse allows ~6e-5."""
    x = np.array([[1.5, 2], [1, 3.5], [3.5, 4]], dtype=np.float32)
    c = np.array([[1.0], [2.0], [3.0]], dtype=np.float32)
    ops = [{"kind": "shift_bounds", "margin": 0.1, "bounds": ()},
           {"kind": "coupling", "knots": 16, "layers": (128, 128)},
           {"kind": "roll", "shift": 1},
           {"kind": "coupling", "knots": 16, "layers": (128, 128)}]
    v, y, ld = _train_then_eval(ops, x, c)
    x2 = zo.chain_inverse(ops, v, y, c)
    assert_allclose(x2, x, rtol=1e-4)


def test_single_coupling_roundtrip_and_split():
    """tests/test_bijectors.py:229-242."""
    x = np.array([[1.5, 2], [1, 3.5], [3.5, 4]], dtype=np.float32)
    c = np.array([[1.0], [2.0], [3.0]], dtype=np.float32)
    ops = [{"kind": "coupling", "knots": 16, "layers": (128, 128)}]
    v = zo.init_variables(ops, 2, 1, 0)
    y, ld, _ = zo.chain_forward(ops, v, x, c)
    assert_allclose(zo.chain_inverse(ops, v, y, c), x, atol=1e-5)  # all oob ⇒ identity
    x3 = np.zeros((3, 3), np.float32)
    xt, xc, _, _ = zo.coupling_params(
        x3, None, zo.init_variables(ops, 3, 0)["params"]["bijectors_0"],
        zo.init_variables(ops, 3, 0)["batch_stats"]["bijectors_0"], knots_=16, train=False)
    assert xt.shape[1] == 1 and xc.shape[1] == 2


def test_rolling_spline_coupling():
    """tests/test_bijectors.py:245-266."""
    x = np.array([[1.5, 2], [1, 3.5], [3.5, 4]], dtype=np.float32)
    c = np.array([[1.0], [2.0], [3.0]], dtype=np.float32)
    ops = zo.make_chain(2, layers=(64, 64))
    assert [o["kind"] for o in ops] == ["shift_bounds", "coupling", "roll", "coupling"]
    v, y, ld = _train_then_eval(ops, x, c)
    assert_allclose(zo.chain_inverse(ops, v, y, c), x, atol=1e-4)
    with pytest.raises(ValueError):
        zo.make_chain(0)
    with pytest.raises(ValueError):
        zo.make_chain(1)


def test_latent_logpdfs():
    """tests/test_distributions.py:11-74 — closed forms (scipy.stats here, jax.scipy there)."""
    rng = np.random.default_rng(1)
    x = rng.uniform(size=(10, 3))
    assert_allclose(zo.latent_log_prob(np.zeros((10, 3)), "uniform"), 0)
    mvn = sps.multivariate_normal(0.5 * np.ones(3), np.identity(3) * 0.1 ** 2)
    assert_allclose(zo.latent_log_prob(x, "normal"), mvn.logpdf(x), atol=1e-5)
    assert_allclose(zo.latent_log_prob(x, "truncnorm"), mvn.logpdf(x), atol=5e-6)
    assert_allclose(zo.latent_log_prob(x, "truncnorm"),
                    sps.truncnorm.logpdf(x, -5, 5, loc=0.5, scale=0.1).sum(-1), rtol=1e-12)
    assert_allclose(zo.latent_log_prob(x, "beta"), sps.beta.logpdf(x, 12, 12).sum(-1))
    assert_allclose(-zo._betaln(12, 12), 16.602059876, rtol=1e-9)
    x32 = x.astype(np.float32)
    assert_allclose(zo.latent_log_prob(x32, "beta"), sps.beta.logpdf(x, 12, 12).sum(-1), rtol=2e-5)
    out = np.array([[-0.1, 0.5], [0.5, 1.5]])
    for kind in ("beta", "uniform", "truncnorm"):
        assert np.isneginf(zo.latent_log_prob(out, kind)).all()


def test_flow_shapes_and_nan_to_num():
    """tests/test_flow.py:7-29 and flow.py:47."""
    x = np.array([[3.0, 2.0], [1.0, 4.0], [5.0, 6.0]], dtype=np.float32)
    ops = [{"kind": "shift_bounds", "margin": 0.1, "bounds": ()}]
    v = zo.init_variables(ops, 2, 0)
    lp, stats = zo.flow_log_prob(ops, v, x, train=True)
    assert lp.shape == (3,) and np.isfinite(lp).all()
    u = np.random.default_rng(0).beta(12, 12, size=(1000, 2)).astype(np.float32)
    x2 = zo.flow_inverse(ops, {"batch_stats": stats}, u)
    assert x2.shape == (1000, 2)
    assert x2[:, 0].min() >= 1 - 0.2 and x2[:, 0].max() <= 5 + 0.2
    lp = zo.nan_to_num(np.array([np.nan, np.inf, -np.inf, 1.0], np.float32))
    fi = np.finfo(np.float32)
    assert_array_equal(lp, np.array([fi.min, fi.max, fi.min, 1.0], np.float32))


def test_idx_equals_K_semantics():
    """an in-range x at/after the last knot gets idx == K and NaN (fill-mode
    gather); an out-of-range x is returned unchanged with zero log-det contribution."""
    K = 4
    dx = np.full((2, 1, K), 0.2499999, np.float32)  # cumsum total < 1
    dy = np.full((2, 1, K), 0.25, np.float32)
    sl = np.ones((2, 1, K - 1), np.float32)
    x = np.array([[0.99999994], [1.5]], np.float32)
    y, ld, idx = zo.rqs_forward(x, dx, dy, sl, return_idx=True)
    assert idx[0, 0] == K and np.isnan(y[0, 0]) and np.isnan(ld[0])
    assert idx[1, 0] == K and y[1, 0] == np.float32(1.5) and ld[1] == 0


def test_torch_oracle_matches_numpy():
    """The gradient oracle's forward (torch float64) equals the numpy oracle's train-mode
    forward in float64, and its gradients agree with central differences of the numpy oracle
    (the reference's own method for derivatives, tests/test_utils.py:38-47)."""
    from oracle import torch_oracle as to

    rng = np.random.default_rng(0)
    D, C, M = 4, 2, 64
    ops = zo.make_chain(D, 8, (16, 16))
    x = rng.normal(0.3, 1.0, (M, D))
    c = rng.uniform(0, 1, (M, C))
    v = zo.init_variables(ops, D, C, 1, weight_scale=2.0, randomize_bn=True)
    v64 = {"params": _to64(v["params"]), "batch_stats": _to64(v["batch_stats"])}
    lp_np, stats_np = zo.flow_log_prob(ops, v64, x, c, train=True)
    loss, grads, new_stats, gc, lp_t = to.loss_and_grads(ops, v64, x, c)
    assert_allclose(lp_t, lp_np, rtol=1e-10, atol=1e-10)
    assert_allclose(loss, -lp_np.mean(), rtol=1e-12)
    assert_allclose(new_stats["bijectors_1"]["BatchNorm_0"]["mean"], stats_np["bijectors_1"]["BatchNorm_0"]["mean"], rtol=1e-6)  # stored as float32 leaves
    assert_allclose(new_stats["bijectors_0"]["xmin_0"], stats_np["bijectors_0"]["xmin_0"], rtol=1e-6)

    def loss_np(vv, cc):
        lp, _ = zo.flow_log_prob(ops, vv, x, cc, train=True)
        return -lp.mean()

    h = 1e-6
    for (name, layer, leaf, pos) in [("bijectors_1", "Dense_2", "kernel", (3, 5)), ("bijectors_3", "Dense_0", "kernel", (1, 2)),
                                     ("bijectors_5", "BatchNorm_0", "scale", (2,)), ("bijectors_1", "Dense_1", "bias", (7,))]:
        vp, vm = _to64(v64), _to64(v64)
        vp["params"][name][layer][leaf][pos] += h
        vm["params"][name][layer][leaf][pos] -= h
        num = (loss_np(vp, c) - loss_np(vm, c)) / (2 * h)
        assert_allclose(grads[name][layer][leaf][pos], num, rtol=2e-5, atol=1e-8)
    cp, cm = c.copy(), c.copy()
    cp[5, 1] += h
    cm[5, 1] -= h
    assert_allclose(gc[5, 1], (loss_np(v64, cp) - loss_np(v64, cm)) / (2 * h), rtol=2e-5, atol=1e-9)


def _to64(t):
    if isinstance(t, dict):
        return {k: _to64(v) for k, v in t.items()}
    return np.array(t, np.float64)
