"""Spline-stage operators on raw conditioner output (CUDA only).

Host-side mirror of the reference's ``zenflow/utils.py`` for the hot path: the reference
calls ``normalize_spline_params`` (utils.py:37-62) and then
``rational_quadratic_spline_forward`` / ``_inverse`` (utils.py:65-202); here both steps are
one fused kernel that reads the raw ``(M, d, 3K-1)`` parameters once from HBM.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._device import like_input, ptr, stream_ptr, to_device_f32

EPS = 1e-5  # utils.py:15

__all__ = ["EPS", "rqs_forward_raw", "rqs_inverse_raw"]


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
