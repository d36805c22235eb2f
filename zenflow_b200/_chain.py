"""Host-side assembly of the native chain description (zf_chain) and the three eval calls.

A ``ChainSpec`` is what the bijector modules emit: the op list of a Chain with device
pointers to the FLAX variable leaves.  It owns the torch tensors behind those pointers for
the duration of the call.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._device import ptr, require_cuda, stream_ptr, to_device_f32


class ChainSpec:
    def __init__(self, dim: int, cdim: int):
        if dim > _lib.ZF_MAX_DIM:
            raise ValueError(f"at most {_lib.ZF_MAX_DIM} columns are supported (got {dim})")
        self.dim = int(dim)
        self.cdim = int(cdim)
        self.device = require_cuda()
        self._ops: List[_lib.ZfOp] = []
        self._keep: list = []  # tensors and ctypes structs that must outlive the call

    # -- emitters used by the bijectors -------------------------------------------------------
    def leaf(self, a) -> torch.Tensor:
        t = to_device_f32(a, self.device)
        self._keep.append(t)
        return t

    def add_roll(self, shift: int) -> None:
        op = _lib.ZfOp()
        op.kind = _lib.OP_ROLL
        op.shift = int(shift)
        self._ops.append(op)

    def add_shift_bounds(self, kinds: Sequence[int], lo: Sequence[float], hi: Sequence[float], margin: float,
                         xmin: Optional[torch.Tensor], xmax: Optional[torch.Tensor]) -> None:
        sb = _lib.ZfShiftBounds()
        for i in range(self.dim):
            sb.kind[i] = int(kinds[i])
            sb.lo[i] = float(lo[i])
            sb.hi[i] = float(hi[i])
        sb.margin = float(margin)
        sb.xmin = ptr(xmin)
        sb.xmax = ptr(xmax)
        self._keep += [sb, xmin, xmax]
        op = _lib.ZfOp()
        op.kind = _lib.OP_SHIFT_BOUNDS
        op.shift_bounds = C.pointer(sb)
        self._ops.append(op)

    def add_coupling(self, knots: int, hidden: Sequence[int], bn_scale, bn_bias, bn_mean, bn_var,
                     kernels: Sequence, biases: Sequence, act: int = 0) -> None:
        if len(hidden) > _lib.ZF_MAX_LAYERS:
            raise ValueError(f"at most {_lib.ZF_MAX_LAYERS} hidden layers are supported")
        cp = _lib.ZfCoupling()
        cp.knots = int(knots)
        cp.n_hidden = len(hidden)
        cp.act = int(act)
        for i, w in enumerate(hidden):
            cp.hidden[i] = int(w)
        cp.bn_scale = ptr(self.leaf(bn_scale))
        cp.bn_bias = ptr(self.leaf(bn_bias))
        cp.bn_mean = ptr(self.leaf(bn_mean))
        cp.bn_var = ptr(self.leaf(bn_var))
        d = self.dim // 2
        fan_in = self.dim - d + self.cdim
        widths = list(hidden) + [d * (3 * knots - 1)]
        for i, (k, b) in enumerate(zip(kernels, biases)):
            kt, bt = self.leaf(k), self.leaf(b)
            if tuple(kt.shape) != (fan_in, widths[i]) or tuple(bt.shape) != (widths[i],):
                raise ValueError(f"Dense_{i}: kernel {tuple(kt.shape)} / bias {tuple(bt.shape)} do not match "
                                 f"the expected ({fan_in}, {widths[i]})")
            cp.kernel[i] = ptr(kt)
            cp.bias[i] = ptr(bt)
            fan_in = widths[i]
        self._keep.append(cp)
        op = _lib.ZfOp()
        op.kind = _lib.OP_COUPLING
        op.coupling = C.pointer(cp)
        self._ops.append(op)

    # -- native calls -------------------------------------------------------------------------
    def _chain(self) -> _lib.ZfChain:
        n = len(self._ops)
        arr = (_lib.ZfOp * max(n, 1))(*self._ops)
        ch = _lib.ZfChain()
        ch.dim, ch.cdim, ch.n_ops = self.dim, self.cdim, n
        ch.ops = C.cast(arr, C.POINTER(_lib.ZfOp))
        self._keep += [arr, ch]
        return ch

    def _inputs(self, x, c):
        xd = to_device_f32(x, self.device)
        if xd.ndim != 2 or xd.shape[1] != self.dim:
            raise ValueError(f"x must have shape (N, {self.dim}), got {tuple(xd.shape)}")
        cd = None
        if self.cdim:
            if c is None:
                raise ValueError("this chain is conditional: c is required")
            cd = to_device_f32(c, self.device)
            if cd.ndim == 1:
                cd = cd.reshape(-1, 1)
            if cd.shape != (xd.shape[0], self.cdim):
                raise ValueError(f"c must have shape ({xd.shape[0]}, {self.cdim}), got {tuple(cd.shape)}")
        return xd, cd

    def _workspace(self, lib, ch, M):
        nbytes = int(lib.zf_chain_workspace_bytes(C.byref(ch), M))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        return ws, nbytes

    def forward(self, x, c, want_y: bool = True, want_log_det: bool = True):
        lib = _lib.load()
        xd, cd = self._inputs(x, c)
        M = xd.shape[0]
        ch = self._chain()
        ws, nbytes = self._workspace(lib, ch, M)
        y = torch.empty_like(xd) if want_y else None
        ld = torch.empty(M, dtype=torch.float32, device=self.device) if want_log_det else None
        _lib.check(lib.zf_chain_forward(stream_ptr(), C.byref(ch), ptr(xd), ptr(cd), M, ptr(y), ptr(ld),
                                        ptr(ws), nbytes), "zf_chain_forward")
        return y, ld

    def bin_indices(self, x, c):
        """Bin index of every spline evaluation of the forward chain: (M, n_couplings, D//2) int32 in [0, K]
        (zf_chain_bin_indices; parity evidence)."""
        lib = _lib.load()
        xd, cd = self._inputs(x, c)
        M = xd.shape[0]
        ch = self._chain()
        ws, nbytes = self._workspace(lib, ch, M)
        n_c = sum(1 for op in self._ops if op.kind == _lib.OP_COUPLING)
        idx = torch.full((M, n_c, max(1, self.dim // 2)), -1, dtype=torch.int32, device=self.device)
        _lib.check(lib.zf_chain_bin_indices(stream_ptr(), C.byref(ch), ptr(xd), ptr(cd), M, ptr(idx), ptr(ws), nbytes),
                   "zf_chain_bin_indices")
        return idx

    def inverse(self, z, c):
        lib = _lib.load()
        zd, cd = self._inputs(z, c)
        M = zd.shape[0]
        ch = self._chain()
        ws, nbytes = self._workspace(lib, ch, M)
        x = torch.empty_like(zd)
        _lib.check(lib.zf_chain_inverse(stream_ptr(), C.byref(ch), ptr(zd), ptr(cd), M, ptr(x), ptr(ws), nbytes),
                   "zf_chain_inverse")
        return x

    def log_prob(self, x, c, latent_kind: int, peakness: float):
        lib = _lib.load()
        xd, cd = self._inputs(x, c)
        M = xd.shape[0]
        ch = self._chain()
        ws, nbytes = self._workspace(lib, ch, M)
        lp = torch.empty(M, dtype=torch.float32, device=self.device)
        _lib.check(lib.zf_flow_log_prob(stream_ptr(), C.byref(ch), int(latent_kind), float(peakness), ptr(xd),
                                        ptr(cd), M, ptr(lp), ptr(ws), nbytes), "zf_flow_log_prob")
        return lp

    def sample(self, n: int, c, latent_kind: int, peakness: float, seed: int):
        """Flow.sample: latent draw inside the inverse chain kernel (zf_flow_sample)."""
        lib = _lib.load()
        cd = None
        if self.cdim:
            if c is None:
                raise ValueError("this chain is conditional: pass one condition vector per sample")
            cd = to_device_f32(c, self.device)
            if cd.ndim == 1:
                cd = cd.reshape(-1, 1)
            if cd.shape != (n, self.cdim):
                raise ValueError(f"c must have shape ({n}, {self.cdim}), got {tuple(cd.shape)}")
        ch = self._chain()
        ws, nbytes = self._workspace(lib, ch, n)
        x = torch.empty((n, self.dim), dtype=torch.float32, device=self.device)
        _lib.check(lib.zf_flow_sample(stream_ptr(), C.byref(ch), int(latent_kind), float(peakness),
                                      int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(cd), n, ptr(x), ptr(ws), nbytes),
                   "zf_flow_sample")
        return x
