"""GPU parity of the standalone spline stage (zf_rqs_forward / zf_rqs_inverse) against the
oracle on identical raw parameters: bin indices bit-exact, values within fp32 tolerance."""
import numpy as np
import pytest
import torch

from oracle import zenflow_oracle as zo
from tests.helpers import assert_fp32_parity

pytestmark = pytest.mark.gpu

Y_ATOL = 2e-6      # y, x in [0, 1]
LD_RTOL = 1e-5     # log-det: rel 1e-5 (north_star) with an absolute floor per transformed dim
LD_ATOL = 2e-6


def _case(M, d, K, seed, scale=1.5, lo=-0.1, hi=1.1):
    rng = np.random.default_rng(seed)
    theta = (scale * rng.standard_normal((M, d, 3 * K - 1))).astype(np.float32)
    x = rng.uniform(lo, hi, (M, d)).astype(np.float32)
    x.reshape(-1)[:4] = [0.0, 1.0, np.float32(1 - 2 ** -24), np.float32(1e-30)][: min(4, x.size)]
    return theta, x


@pytest.mark.parametrize("M,d,K", [(1000, 1, 16), (4099, 8, 32), (777, 3, 5), (513, 2, 3), (64, 1, 1),
                                   (300_001, 1, 16), (70_003, 8, 32), (2050, 5, 16), (129, 40, 4)])
def test_forward_matches_oracle(M, d, K):
    from zenflow_b200.utils import rqs_forward_raw

    theta, x = _case(M, d, K, seed=M + d + K)
    y, ld, idx = rqs_forward_raw(x, theta, K, return_index=True)
    yo, ldo, idxo = zo.rqs_forward_theta(x, theta, K, return_idx=True)
    assert idx.dtype == np.int32
    np.testing.assert_array_equal(idx, idxo)  # bit-exact bins
    y64, ld64 = zo.rqs_forward_theta(x.astype(np.float64), theta.astype(np.float64), K)
    good = np.isfinite(yo)
    np.testing.assert_array_equal(np.isnan(y), np.isnan(yo))
    np.testing.assert_allclose(y[good], yo[good], atol=Y_ATOL, rtol=0)
    # vs float64 truth: as good as the reference's float32 arithmetic (see tests/helpers.py)
    assert_fp32_parity(np.where(good, y, 0), np.where(good, y64, 0), np.where(good, yo, 0), "y", rtol=0, atol=2 * Y_ATOL)
    goodl = np.isfinite(ldo)
    np.testing.assert_allclose(ld[goodl], ldo[goodl], rtol=LD_RTOL, atol=LD_ATOL * d)
    assert_fp32_parity(ld[goodl], ld64[goodl], ldo[goodl], "log_det", rtol=LD_RTOL, atol=2 * LD_ATOL * d)
    oob = (x < 0) | (x >= 1)
    np.testing.assert_array_equal(y[oob], x[oob])  # identity outside [0, 1) is exact


@pytest.mark.parametrize("M,d,K", [(1000, 1, 16), (4099, 8, 32), (777, 3, 5), (513, 2, 3), (300_001, 1, 16),
                                   (70_003, 8, 32)])
def test_inverse_matches_oracle(M, d, K):
    from zenflow_b200.utils import rqs_inverse_raw

    theta, y = _case(M, d, K, seed=7 * M + d + K)
    x, idx = rqs_inverse_raw(y, theta, K, return_index=True)
    xo, idxo = zo.rqs_inverse_theta(y, theta, K, return_idx=True)
    np.testing.assert_array_equal(idx, idxo)
    good = np.isfinite(xo)
    np.testing.assert_array_equal(np.isnan(x), np.isnan(xo))
    np.testing.assert_allclose(x[good], xo[good], atol=5e-6, rtol=0)
    oob = (y < 0) | (y >= 1)
    np.testing.assert_array_equal(x[oob], y[oob])


def test_extreme_parameters_take_the_ieee_path():
    """|theta| large / quotients tiny: the per-row range check must route to the IEEE path and
    still agree bit-for-bit on the bins."""
    from zenflow_b200.utils import rqs_forward_raw

    K, d, M = 16, 1, 4096
    rng = np.random.default_rng(5)
    theta = (2.0 * rng.standard_normal((M, d, 3 * K - 1))).astype(np.float32)
    theta[::3, 0, 1] = -1e6      # squareplus underflows to 0
    theta[1::3, 0, 5] = 3e7      # one bin takes everything
    theta[2::7, 0, 20] = -4e8
    x = rng.uniform(0, 1, (M, d)).astype(np.float32)
    y, ld, idx = rqs_forward_raw(x, theta, K, return_index=True)
    yo, ldo, idxo = zo.rqs_forward_theta(x, theta, K, return_idx=True)
    np.testing.assert_array_equal(idx, idxo)
    good = np.isfinite(yo) & np.isfinite(y)
    np.testing.assert_allclose(y[good], yo[good], atol=5e-6)


@pytest.mark.parametrize("K,d", [(16, 1), (32, 8)])
def test_nan_inf_and_large_parameters_lean_row(K, d):
    """The lean row of the stage kernel (K = 16 / 32) tests every searched / other-axis value against the fast-path bound;
    NaN, inf and |theta| >= 4096 must route the row to the IEEE path: bins bit-exact, NaN pattern as the oracle's, forward
    and inverse."""
    from zenflow_b200.utils import rqs_forward_raw, rqs_inverse_raw

    M = 3000
    rng = np.random.default_rng(K + d)
    theta = (1.5 * rng.standard_normal((M, d, 3 * K - 1))).astype(np.float32)
    theta[::5, 0, 3] = 5000.0                 # just above the bound, searched block (forward)
    theta[1::5, d - 1, K + 2] = -7000.0       # other block (forward) = searched block (inverse)
    theta[2::11, 0, 1] = np.inf
    theta[3::13, d - 1, K + 1] = np.nan
    theta[4::17, 0, 2 * K + 1] = 1e9          # a slope: never decides the path
    x = rng.uniform(-0.05, 1.05, (M, d)).astype(np.float32)
    with np.errstate(all="ignore"):
        yo, ldo, idxo = zo.rqs_forward_theta(x, theta, K, return_idx=True)
        xo, idxi = zo.rqs_inverse_theta(x, theta, K, return_idx=True)
    y, ld, idx = rqs_forward_raw(x, theta, K, return_index=True)
    np.testing.assert_array_equal(idx, idxo)
    np.testing.assert_array_equal(np.isnan(y), np.isnan(yo))
    good = np.isfinite(yo)
    np.testing.assert_allclose(y[good], yo[good], atol=5e-6, rtol=0)
    xi, idx2 = rqs_inverse_raw(x, theta, K, return_index=True)
    np.testing.assert_array_equal(idx2, idxi)
    np.testing.assert_array_equal(np.isnan(xi), np.isnan(xo))
    good = np.isfinite(xo)
    np.testing.assert_allclose(xi[good], xo[good], atol=1e-5, rtol=0)


def test_reference_kats_on_device():
    """tests/test_utils.py:7-13 (identity spline incl. out of range) through the CUDA path:
    raw parameters 0 give equal bins and squareplus(0)=1 slopes."""
    from zenflow_b200.utils import rqs_forward_raw, rqs_inverse_raw

    x = np.linspace(-1, 2, 10).reshape(-1, 1).astype(np.float32)
    theta = np.zeros((10, 1, 11), np.float32)  # K = 4
    y, ld = rqs_forward_raw(x, theta, 4)
    np.testing.assert_allclose(y, x, atol=1e-5)
    x2 = rqs_inverse_raw(y, theta, 4)
    np.testing.assert_allclose(x2, x, atol=1e-4)


def test_torch_in_torch_out_and_unaligned_theta():
    from zenflow_b200.utils import rqs_forward_raw

    theta, x = _case(3000, 1, 16, seed=11)
    big = torch.empty(theta.size + 1, device="cuda")
    tview = big[1:].view(theta.shape)  # 4-byte aligned only: forces the cooperative-load path
    tview.copy_(torch.from_numpy(theta))
    y, ld, idx = rqs_forward_raw(torch.from_numpy(x).cuda(), tview, 16, return_index=True)
    assert isinstance(y, torch.Tensor) and y.is_cuda
    yo, ldo, idxo = zo.rqs_forward_theta(x, theta, 16, return_idx=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), idxo)
    np.testing.assert_allclose(ld.cpu().numpy(), ldo, rtol=LD_RTOL, atol=LD_ATOL)


def test_empty_batch():
    from zenflow_b200.utils import rqs_forward_raw

    y, ld = rqs_forward_raw(np.zeros((0, 2), np.float32), np.zeros((0, 2, 47), np.float32), 16)
    assert y.shape == (0, 2) and ld.shape == (0,)


def test_exact_math_fast_paths_exhaustive():
    """The hand-rolled sqrt.rn / div.rn fast paths that decide bin indices are bit-identical
    to the IEEE instructions over their whole admitted input range."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    bad = torch.zeros(3, dtype=torch.int64, device="cuda")
    _lib.check(lib.zf_selftest_exact_math(torch.cuda.current_stream().cuda_stream, bad.data_ptr()))
    assert bad.tolist() == [0, 0, 0]
