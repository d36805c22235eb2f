"""The reference's own tests of zenflow.utils (tests/test_utils.py) re-run against the CUDA drop-ins of the same
names and signatures, plus oracle parity on random normalised parameters."""
import numpy as np
import pytest

from oracle import zenflow_oracle as zo

pytestmark = pytest.mark.gpu


def test_identity_spline_kat():
    """tests/test_utils.py:7-13: equal bins and unit slopes give the identity on [-1, 2] (atol 1e-5)."""
    from zenflow_b200 import utils

    x = np.linspace(-1, 2, 10, dtype=np.float32).reshape(-1, 1)
    bins = 4
    dx = np.full((10, 1, bins), 1.0 / bins, np.float32)
    dy = np.full((10, 1, bins), 1.0 / bins, np.float32)
    slope = np.ones((10, 1, bins - 1), np.float32)
    y, log_det = utils.rational_quadratic_spline_forward(x, dx, dy, slope)
    np.testing.assert_allclose(y, x, atol=1e-5)
    yo, ldo = zo.rqs_forward(x, dx, dy, slope)
    np.testing.assert_allclose(y, yo, atol=1e-6)
    np.testing.assert_allclose(log_det, ldo, atol=1e-6)


def test_log_det_is_log_of_numeric_derivative_and_inverse_round_trip():
    """tests/test_utils.py:16-50: log_det == log(dy/dx) (central differences in float64 of the oracle play jacobi's
    role, atol 0.01) on 1000 points in [-0.1, 1.1], K = 3, default_rng(1); inverse(forward(x)) = x (atol 1e-4)."""
    from zenflow_b200 import utils

    rng = np.random.default_rng(1)
    K = 3
    raw = [rng.normal(size=(1000, 1, n)).astype(np.float32) for n in (K, K, K - 1)]
    dx, dy, slope = utils.normalize_spline_params(*raw)
    odx, ody, osl = zo.normalize_spline_params(*raw)
    np.testing.assert_array_equal(dx, odx)       # the knot path is bit-exact
    np.testing.assert_array_equal(dy, ody)
    np.testing.assert_allclose(slope, osl, rtol=1e-7)
    np.testing.assert_allclose(dx.sum(-1), 1, atol=1e-6)   # tests/test_utils.py:77-94 (rows sum to 1, entries >= threshold)
    assert dx.min() >= zo.EPS * (1 - 1e-6)
    x = np.linspace(-0.1, 1.1, 1000, dtype=np.float32).reshape(-1, 1)
    y, log_det = utils.rational_quadratic_spline_forward(x, dx, dy, slope)
    h = 1e-6
    f = lambda t: zo.rqs_forward(t, dx.astype(np.float64), dy.astype(np.float64), slope.astype(np.float64))[0]
    x64 = x.astype(np.float64)
    num = (f(x64 + h) - f(x64 - h))[:, 0] / (2 * h)
    keep = np.abs(num) > 0
    np.testing.assert_allclose(log_det[keep], np.log(num[keep]), atol=0.01)
    x2 = utils.rational_quadratic_spline_inverse(y, dx, dy, slope)
    np.testing.assert_allclose(x2, x, atol=1e-4)


def test_squareplus_and_random_parity():
    import torch

    from zenflow_b200 import utils

    rng = np.random.default_rng(3)
    a = rng.normal(scale=3, size=10_001).astype(np.float32)
    np.testing.assert_array_equal(utils.squareplus(a), zo.squareplus(a))
    M, d, K = 3001, 3, 8
    raw = [rng.normal(size=(M, d, n)).astype(np.float32) for n in (K, K, K - 1)]
    dx, dy, slope = zo.normalize_spline_params(*raw)
    x = rng.uniform(-0.05, 1.05, (M, d)).astype(np.float32)
    y, ld = utils.rational_quadratic_spline_forward(x, dx, dy, slope)
    yo, ldo, idxo = zo.rqs_forward(x, dx, dy, slope, return_idx=True)
    y64, ld64 = zo.rqs_forward(x.astype(np.float64), dx.astype(np.float64), dy.astype(np.float64), slope.astype(np.float64))
    assert np.abs(y - y64).max() <= 2 * np.abs(yo - y64).max() + 1e-6
    assert np.abs(ld - ld64).max() <= 2 * np.abs(ldo - ld64).max() + 1e-5
    xi = utils.rational_quadratic_spline_inverse(y, dx, dy, slope)
    xio = zo.rqs_inverse(y, dx, dy, slope)
    np.testing.assert_allclose(xi, xio, atol=1e-5)
    # torch in -> torch out
    yt, _ = utils.rational_quadratic_spline_forward(torch.from_numpy(x).cuda(), torch.from_numpy(dx).cuda(),
                                                    torch.from_numpy(dy).cuda(), torch.from_numpy(slope).cuda())
    assert isinstance(yt, torch.Tensor) and yt.is_cuda
    with pytest.raises(ValueError):
        utils.rational_quadratic_spline_forward(x, dx, dy[:, :, :-1], slope)
