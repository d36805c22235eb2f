"""Deep-Set conditioner of the reference's deep_set example, on the native path.

Host-side mirror of the user modules in examples/deep_set.ipynb (code cells 3 and 5): ``Phi`` (BatchNorm ->
NNBlock(8, 3, 128) -> Dropout(0.3) -> sum-pool with the BCOO matrix) and ``DeepSetFlow`` (Phi feeding the
conditions of a spline Flow, trained jointly through d loss / d c).  The arithmetic is zf_phi_forward /
zf_phi_backward (csrc/zf_phi.cu) and zf_flow_value_and_grad; nothing is computed in Python.

The notebook's ``sum_matrix`` (a BCOO matrix of ones) is passed as ``SumMatrix(set_idx, row_idx, n_sets)``, its COO
index list (``SumMatrix.from_sizes`` builds what the notebook's ``preprocess`` builds).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._device import like_input, ptr, require_cuda, stream_ptr, to_device_f32

__all__ = ["Phi", "SumMatrix", "PhiEngine", "DeepSetFlowTrainer"]


@dataclass
class SumMatrix:
    """COO form of the notebook's BCOO sum matrix of ones: c[set_idx[e]] += h[row_idx[e]]."""

    set_idx: torch.Tensor
    row_idx: torch.Tensor
    n_sets: int

    def __post_init__(self):
        # the kernels index with these: validate once here (an out-of-range entry would write outside c / read outside h)
        dev = require_cuda()
        as_i32 = lambda t: torch.as_tensor(t).to(device=dev, dtype=torch.int32).contiguous().reshape(-1)
        self.set_idx, self.row_idx, self.n_sets = as_i32(self.set_idx), as_i32(self.row_idx), int(self.n_sets)
        if self.set_idx.numel() != self.row_idx.numel():
            raise ValueError(f"SumMatrix: {self.set_idx.numel()} set indices but {self.row_idx.numel()} row indices")
        if self.n_sets < 1:
            raise ValueError("SumMatrix: n_sets must be at least 1")
        self._max_row = -1
        if self.set_idx.numel():
            s_lo, s_hi = (int(v) for v in torch.aminmax(self.set_idx))
            r_lo, r_hi = (int(v) for v in torch.aminmax(self.row_idx))
            if s_lo < 0 or s_hi >= self.n_sets:
                raise ValueError(f"SumMatrix: set indices must be in [0, {self.n_sets}) (found {s_lo} .. {s_hi})")
            if r_lo < 0:
                raise ValueError(f"SumMatrix: negative row index {r_lo}")
            self._max_row = r_hi

    def check_rows(self, n_rows: int) -> None:
        if self._max_row >= n_rows:
            raise ValueError(f"SumMatrix: row index {self._max_row} is out of bounds for {n_rows} rows")

    @staticmethod
    def from_sizes(sizes: Sequence[int], device=None) -> "SumMatrix":
        device = device or require_cuda()
        sizes = np.asarray(sizes, dtype=np.int64)
        set_idx = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
        row_idx = np.arange(int(sizes.sum()), dtype=np.int32)
        return SumMatrix(torch.from_numpy(set_idx).to(device), torch.from_numpy(row_idx).to(device), int(len(sizes)))

    @property
    def nnz(self) -> int:
        return int(self.set_idx.numel())


class Phi:
    """deep_set.ipynb:152-160.  Variable tree as FLAX names it: params/{BatchNorm_0/{scale,bias},
    NNBlock_0/Dense_j/{kernel,bias}}, batch_stats/BatchNorm_0/{mean,var}."""

    def __init__(self, out_dim: int = 8, depth: int = 3, width: int = 128, rate: float = 0.3):
        self.out_dim, self.depth, self.width, self.rate = int(out_dim), int(depth), int(width), float(rate)

    def init(self, seed: int, x) -> Dict[str, dict]:
        in_dim = int(np.asarray(x.shape)[-1])
        rng = np.random.default_rng(seed)
        block = {}
        fan_in = in_dim
        for j, w in enumerate([self.width] * self.depth + [self.out_dim]):
            block[f"Dense_{j}"] = {"kernel": (rng.standard_normal((fan_in, w)) / np.sqrt(fan_in)).astype(np.float32),
                                   "bias": np.zeros(w, np.float32)}
            fan_in = w
        return {"params": {"BatchNorm_0": {"scale": np.ones(in_dim, np.float32), "bias": np.zeros(in_dim, np.float32)},
                           "NNBlock_0": block},
                "batch_stats": {"BatchNorm_0": {"mean": np.zeros(in_dim, np.float32), "var": np.ones(in_dim, np.float32)}}}

    def apply(self, variables, x, sum_matrix: SumMatrix, train: bool = False, *, seed: int = 0, dropout_mask=None):
        """c = Phi(x, sum_matrix, train).  Eval returns c; train returns (c, {"batch_stats": ...}) like
        ``apply(..., mutable=["batch_stats"])``."""
        eng = PhiEngine(self, variables, int(np.asarray(x.shape)[-1]))
        c = eng.forward(x, sum_matrix, train=train, seed=seed, dropout_mask=dropout_mask)
        c = like_input(c, x)
        if not train:
            return c
        return c, {"batch_stats": eng.variables(as_numpy=not isinstance(x, torch.Tensor))["batch_stats"]}


class PhiEngine:
    """Device-resident Phi: flat parameter / gradient buffers, forward and backward through the C ABI."""

    def __init__(self, phi: Phi, variables, in_dim: int):
        self.phi, self.dev, self.in_dim = phi, require_cuda(), int(in_dim)
        p, st = variables["params"], variables["batch_stats"]
        names = [("BatchNorm_0", "scale"), ("BatchNorm_0", "bias")]
        n_dense = phi.depth + 1
        for j in range(n_dense):
            names += [(f"Dense_{j}", "kernel"), (f"Dense_{j}", "bias")]
        get = lambda m, l: p["BatchNorm_0"][l] if m == "BatchNorm_0" else p["NNBlock_0"][m][l]
        total = sum(int(np.prod(get(m, l).shape)) for m, l in names)
        self.n_params = total
        self.P = torch.empty(total, dtype=torch.float32, device=self.dev)
        self.G = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.pv, self.gv = {}, {}
        off = 0
        for m, l in names:
            src = get(m, l)
            n = int(np.prod(src.shape))
            self.pv[(m, l)] = self.P[off:off + n].view(tuple(src.shape))
            self.pv[(m, l)].copy_(to_device_f32(src, self.dev))
            self.gv[(m, l)] = self.G[off:off + n].view(tuple(src.shape))
            off += n
        self.ra_mean = to_device_f32(st["BatchNorm_0"]["mean"], self.dev).clone()
        self.ra_var = to_device_f32(st["BatchNorm_0"]["var"], self.dev).clone()
        z = _lib.ZfPhi()
        z.in_dim, z.out_dim, z.n_hidden = self.in_dim, phi.out_dim, phi.depth
        for i in range(phi.depth):
            z.hidden[i] = phi.width
        z.bn_scale, z.bn_bias = ptr(self.pv[("BatchNorm_0", "scale")]), ptr(self.pv[("BatchNorm_0", "bias")])
        z.bn_mean, z.bn_var = ptr(self.ra_mean), ptr(self.ra_var)
        g = _lib.ZfCouplingGrads()
        g.bn_scale, g.bn_bias = ptr(self.gv[("BatchNorm_0", "scale")]), ptr(self.gv[("BatchNorm_0", "bias")])
        for j in range(n_dense):
            z.kernel[j], z.bias[j] = ptr(self.pv[(f"Dense_{j}", "kernel")]), ptr(self.pv[(f"Dense_{j}", "bias")])
            g.kernel[j], g.bias[j] = ptr(self.gv[(f"Dense_{j}", "kernel")]), ptr(self.gv[(f"Dense_{j}", "bias")])
        self.z, self.g = z, g
        self._ws: Optional[torch.Tensor] = None
        self._fwd = None   # arguments of the last train-mode forward (the backward must repeat them)

    def _workspace(self, N: int) -> torch.Tensor:
        need = int(_lib.load().zf_phi_workspace_bytes(C.byref(self.z), N))
        if need == 0:
            _lib.check(1, "zf_phi_workspace_bytes")
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
        return self._ws

    def forward(self, x, sm: SumMatrix, *, train: bool = False, seed: int = 0, dropout_mask=None) -> torch.Tensor:
        lib = _lib.load()
        x = to_device_f32(x, self.dev)
        N = x.shape[0]
        sm.check_rows(N)
        ws = self._workspace(N)
        base = (ws.data_ptr() + 255) & ~255
        mask = None if dropout_mask is None else to_device_f32(dropout_mask, self.dev)
        c = torch.empty(sm.n_sets, self.phi.out_dim, dtype=torch.float32, device=self.dev)
        _lib.check(lib.zf_phi_forward(stream_ptr(), C.byref(self.z), ptr(x), N, ptr(sm.set_idx), ptr(sm.row_idx), sm.nnz,
                                      sm.n_sets, int(bool(train)), self.phi.rate, int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(mask),
                                      ptr(c), base, ws.numel() - (base - ws.data_ptr())), "zf_phi_forward")
        self._fwd = (x, sm, int(seed) & 0xFFFFFFFFFFFFFFFF, mask) if train else None
        return c

    def backward(self, gc: torch.Tensor) -> None:
        """Parameter gradients (+= into G) of the last train-mode forward for d loss / d c = gc."""
        if self._fwd is None:
            raise RuntimeError("PhiEngine.backward needs a preceding forward(train=True)")
        lib = _lib.load()
        x, sm, seed, mask = self._fwd
        ws = self._ws
        base = (ws.data_ptr() + 255) & ~255
        gc = to_device_f32(gc, self.dev)
        _lib.check(lib.zf_phi_backward(stream_ptr(), C.byref(self.z), C.byref(self.g), ptr(x), x.shape[0], ptr(sm.set_idx),
                                       ptr(sm.row_idx), sm.nnz, sm.n_sets, self.phi.rate, seed, ptr(mask), ptr(gc), base,
                                       ws.numel() - (base - ws.data_ptr())), "zf_phi_backward")

    def variables(self, as_numpy: bool = False) -> Dict[str, dict]:
        conv = (lambda t: t.detach().cpu().numpy().copy()) if as_numpy else (lambda t: t)
        block = {f"Dense_{j}": {"kernel": conv(self.pv[(f"Dense_{j}", "kernel")]), "bias": conv(self.pv[(f"Dense_{j}", "bias")])}
                 for j in range(self.phi.depth + 1)}
        return {"params": {"BatchNorm_0": {"scale": conv(self.pv[("BatchNorm_0", "scale")]),
                                           "bias": conv(self.pv[("BatchNorm_0", "bias")])}, "NNBlock_0": block},
                "batch_stats": {"BatchNorm_0": {"mean": conv(self.ra_mean), "var": conv(self.ra_var)}}}

    def gradients(self) -> Dict[str, dict]:
        block = {f"Dense_{j}": {"kernel": self.gv[(f"Dense_{j}", "kernel")], "bias": self.gv[(f"Dense_{j}", "bias")]}
                 for j in range(self.phi.depth + 1)}
        return {"BatchNorm_0": {"scale": self.gv[("BatchNorm_0", "scale")], "bias": self.gv[("BatchNorm_0", "bias")]},
                "NNBlock_0": block}


class DeepSetFlowTrainer:
    """The jitted ``step`` of deep_set.ipynb:320-352 (DeepSetFlow: c = Phi(x, sum_matrix); loss = -mean(flow(y, c))):
    Phi forward (train) -> zf_flow_value_and_grad (returns d loss / d c) -> Phi backward -> AdamW on both."""

    def __init__(self, phi: Phi, phi_variables, flow, flow_variables, x_dim: int, y_dim: int, *, lr: float = 1e-3,
                 weight_decay: float = 1e-4):
        from ._train import TrainEngine

        self.phi_eng = PhiEngine(phi, phi_variables, x_dim)
        self.flow_eng = TrainEngine(flow, flow_variables, y_dim, phi.out_dim, lr=lr, weight_decay=weight_decay, nesterov=False)
        self.mu = torch.zeros_like(self.phi_eng.P)
        self.nu = torch.zeros_like(self.phi_eng.P)
        self.count = 0
        self.hp = dict(lr=lr, b1=0.9, b2=0.999, eps=1e-8, weight_decay=weight_decay)   # optax.adamw defaults

    def step(self, x, sm: SumMatrix, y, *, seed: int = 0, dropout_mask=None) -> torch.Tensor:
        """One optimiser step; returns the device scalar sum of log-probs (loss = -sum / len(y))."""
        lib = _lib.load()
        pe, fe = self.phi_eng, self.flow_eng
        c = pe.forward(x, sm, train=True, seed=seed, dropout_mask=dropout_mask)
        lp_sum, gc = fe.step(y, c, want_gc=True)
        pe.G.zero_()
        pe.backward(gc)
        h = self.hp
        _lib.check(lib.zf_nadamw_update(stream_ptr(), pe.n_params, ptr(pe.P), ptr(pe.G), ptr(self.mu), ptr(self.nu), self.count,
                                        h["lr"], h["b1"], h["b2"], h["eps"], h["weight_decay"], 0), "zf_nadamw_update")
        self.count += 1
        return lp_sum

    def log_prob(self, x, sm: SumMatrix, y) -> torch.Tensor:
        """metric path (deep_set.ipynb:337-340): eval-mode Phi and flow."""
        c = self.phi_eng.forward(x, sm, train=False)
        return self.flow_eng.flow.apply(self.flow_eng.variables(), to_device_f32(y, self.phi_eng.dev), c)
