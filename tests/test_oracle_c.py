"""The plain-C restatement of the spline stage agrees with the numpy oracle: bins bit for bit,
values to the last ulp or two (libm vs numpy log).  Two independent restatements of the reference's
operation order pin what the CUDA kernels must reproduce."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import zenflow_oracle as zo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "librqs_oracle.so"))
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    lib.zo_rqs_forward.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp, fp, ip]
    lib.zo_rqs_inverse.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp, ip]
    return lib


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t))


@pytest.mark.parametrize("M,d,K", [(5000, 1, 16), (3000, 8, 32), (2000, 3, 5), (1000, 2, 1)])
def test_c_and_numpy_oracles_agree(clib, M, d, K):
    rng = np.random.default_rng(M + K)
    theta = (1.5 * rng.standard_normal((M, d, 3 * K - 1))).astype(np.float32)
    x = rng.uniform(-0.1, 1.1, (M, d)).astype(np.float32)
    x.reshape(-1)[:3] = [0.0, 1.0, np.float32(1 - 2 ** -24)]
    y, ld, idx = np.empty_like(x), np.empty(M, np.float32), np.empty((M, d), np.int32)
    clib.zo_rqs_forward(_p(theta), _p(x), M, d, K, _p(y), _p(ld), _p(idx, ctypes.c_int32))
    yo, ldo, idxo = zo.rqs_forward_theta(x, theta, K, return_idx=True)
    np.testing.assert_array_equal(idx, idxo)
    good = np.isfinite(yo)
    np.testing.assert_array_equal(np.isnan(y), np.isnan(yo))
    np.testing.assert_allclose(y[good], yo[good], rtol=0, atol=3e-7)
    gl = np.isfinite(ldo)
    np.testing.assert_allclose(ld[gl], ldo[gl], rtol=2e-6, atol=2e-6 * d)
    xi, idxi = np.empty_like(x), np.empty((M, d), np.int32)
    clib.zo_rqs_inverse(_p(theta), _p(x), M, d, K, _p(xi), _p(idxi, ctypes.c_int32))
    xo, idxio = zo.rqs_inverse_theta(x, theta, K, return_idx=True)
    np.testing.assert_array_equal(idxi, idxio)
    g = np.isfinite(xo)
    np.testing.assert_allclose(xi[g], xo[g], rtol=0, atol=2e-6)
