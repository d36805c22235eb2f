"""Spline-stage operators on raw conditioner output (CUDA only).

Host-side mirror of the reference's ``zenflow/utils.py``.  The reference calls
``normalize_spline_params`` (utils.py:37-62) and then ``rational_quadratic_spline_forward`` /
``_inverse`` (utils.py:65-202); on the hot path both steps are one fused kernel that reads the raw
``(M, d, 3K-1)`` parameters once from HBM (``rqs_forward_raw`` / ``rqs_inverse_raw``).  The reference's
own public names and signatures (``squareplus``, ``normalize_spline_params``,
``rational_quadratic_spline_forward(x, dx, dy, slope)``, ``..._inverse``) are provided on top of
their own small kernels (zf_utils.cu) so that code written against ``zenflow.utils`` runs unchanged.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._device import like_input, ptr, stream_ptr, to_device_f32

EPS = 1e-5  # utils.py:15

__all__ = ["EPS", "squareplus", "normalize_spline_params", "rational_quadratic_spline_forward",
           "rational_quadratic_spline_inverse", "rqs_forward_raw", "rqs_inverse_raw"]


def _check_shapes(x: torch.Tensor, theta: torch.Tensor, knots: int) -> Tuple[int, int]:
    if x.ndim != 2 or theta.ndim != 3:
        raise ValueError("x must be (M, d) and theta (M, d, 3K-1)")
    M, d = x.shape
    if theta.shape != (M, d, 3 * knots - 1):
        raise ValueError(f"theta has shape {tuple(theta.shape)}, expected {(M, d, 3 * knots - 1)}")
    return M, d


def rqs_forward_raw(x, theta, knots: int, *, return_index: bool = False):
    """normalize_spline_params + rational_quadratic_spline_forward (utils.py:37-141).

    x (M, d), theta (M, d, 3K-1) = raw widths | heights | slopes  ->  y (M, d), log_det (M,)
    [, idx (M, d) int32 bin indices in [0, K]].
    """
    xd, td = to_device_f32(x), to_device_f32(theta)
    M, d = _check_shapes(xd, td, knots)
    y = torch.empty_like(xd)
    ld = torch.empty(M, dtype=torch.float32, device=xd.device)
    idx = torch.empty((M, d), dtype=torch.int32, device=xd.device) if return_index else None
    lib = _lib.load()
    _lib.check(lib.zf_rqs_forward(stream_ptr(), ptr(td), ptr(xd), M, d, knots, ptr(y), ptr(ld), ptr(idx)),
               "zf_rqs_forward")
    out = (like_input(y, x), like_input(ld, x))
    return out + (like_input(idx, x),) if return_index else out


def rqs_inverse_raw(y, theta, knots: int, *, return_index: bool = False):
    """normalize_spline_params + rational_quadratic_spline_inverse (utils.py:144-202)."""
    yd, td = to_device_f32(y), to_device_f32(theta)
    M, d = _check_shapes(yd, td, knots)
    x = torch.empty_like(yd)
    idx = torch.empty((M, d), dtype=torch.int32, device=yd.device) if return_index else None
    lib = _lib.load()
    _lib.check(lib.zf_rqs_inverse(stream_ptr(), ptr(td), ptr(yd), M, d, knots, ptr(x), ptr(idx)),
               "zf_rqs_inverse")
    return (like_input(x, y), like_input(idx, y)) if return_index else like_input(x, y)


# ---------------------------------------------------------------------------------------------
# the reference's public functions on normalised parameters (utils.py:18-62, 65-202)
# ---------------------------------------------------------------------------------------------
def squareplus(x):
    """utils.py:18-20: 0.5 * (x + sqrt(x*x + 4)), elementwise."""
    xd = to_device_f32(x)
    y = torch.empty_like(xd)
    _lib.check(_lib.load().zf_squareplus(stream_ptr(), ptr(xd), xd.numel(), ptr(y)), "zf_squareplus")
    return like_input(y, x)


def normalize_spline_params(dx, dy, slope):
    """utils.py:37-62: raw widths / heights (..., K) and slopes (..., K-1) -> normalised bin widths and heights
    (each row sums to 1, every entry >= EPS) and positive knot derivatives."""
    dxd, dyd, sld = to_device_f32(dx), to_device_f32(dy), to_device_f32(slope)
    K = dxd.shape[-1]
    if dyd.shape != dxd.shape or sld.shape != dxd.shape[:-1] + (K - 1,):
        raise ValueError("dx, dy must be (..., K) and slope (..., K-1)")
    rows = dxd.numel() // K
    theta = torch.cat([dxd.reshape(rows, K), dyd.reshape(rows, K), sld.reshape(rows, K - 1)], dim=1).contiguous()
    odx, ody, osl = torch.empty_like(dxd), torch.empty_like(dyd), torch.empty_like(sld)
    _lib.check(_lib.load().zf_normalize_spline_params(stream_ptr(), ptr(theta), rows, K, ptr(odx), ptr(ody), ptr(osl)),
               "zf_normalize_spline_params")
    return like_input(odx, dx), like_input(ody, dx), like_input(osl, dx)


def _check_normalized(v: torch.Tensor, dx: torch.Tensor, dy: torch.Tensor, slope: torch.Tensor) -> Tuple[int, int, int]:
    if v.ndim != 2 or dx.ndim != 3:
        raise ValueError("inputs must be (M, N) and dx, dy (M, N, K), slope (M, N, K-1)")
    M, d = v.shape
    K = dx.shape[-1]
    if dx.shape != (M, d, K) or dy.shape != (M, d, K) or slope.shape != (M, d, K - 1):
        raise ValueError(f"dx / dy must be {(M, d, K)} and slope {(M, d, K - 1)}")
    return M, d, K


def rational_quadratic_spline_forward(x, dx, dy, slope):
    """utils.py:65-141: x (M, N) and normalised dx, dy (M, N, K), slope (M, N, K-1) -> (y (M, N), log_det (M,))."""
    xd, dxd, dyd, sld = (to_device_f32(a) for a in (x, dx, dy, slope))
    M, d, K = _check_normalized(xd, dxd, dyd, sld)
    y = torch.empty_like(xd)
    ld = torch.empty(M, dtype=torch.float32, device=xd.device)
    _lib.check(_lib.load().zf_rqs_forward_normalized(stream_ptr(), ptr(xd), ptr(dxd), ptr(dyd), ptr(sld), M, d, K, ptr(y),
                                                     ptr(ld), None), "zf_rqs_forward_normalized")
    return like_input(y, x), like_input(ld, x)


def rational_quadratic_spline_inverse(y, dx, dy, slope):
    """utils.py:144-202: the analytic inverse (one array, as the reference returns despite its annotation)."""
    yd, dxd, dyd, sld = (to_device_f32(a) for a in (y, dx, dy, slope))
    M, d, K = _check_normalized(yd, dxd, dyd, sld)
    x = torch.empty_like(yd)
    _lib.check(_lib.load().zf_rqs_inverse_normalized(stream_ptr(), ptr(yd), ptr(dxd), ptr(dyd), ptr(sld), M, d, K, ptr(x),
                                                     None), "zf_rqs_inverse_normalized")
    return like_input(x, y)
