"""Bijectors of the conditional normalizing flow — host-side mirror of zenflow/bijectors.py.

Same classes, fields, call signatures, variable names and error behaviour as the reference;
the bodies hand a description of themselves (``ChainSpec``) to the CUDA library instead of
computing with jax.numpy.  A ``Chain`` evaluated with ``train=False`` is ONE fused kernel
launch for the whole chain (csrc/zf_chain.cu); ``train=True`` runs the per-bijector
train-mode kernels (batch statistics couple the samples).
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from collections.abc import Sequence as _SequenceABC
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._chain import ChainSpec
from ._device import like_input, to_device_f32
from .module import Module, Scope

__all__ = [
    "Bijector",
    "ShiftBounds",
    "Roll",
    "NeuralSplineCoupling",
    "Chain",
    "chain",
    "rolling_spline_coupling",
]


def _shape2(x) -> Tuple[int, int]:
    shp = tuple(x.shape)
    if len(shp) != 2:
        raise ValueError(f"expected a 2-D array (N, D), got shape {shp}")
    return shp


def _cdim(c) -> int:
    if c is None:
        return 0
    return 1 if len(c.shape) == 1 else int(c.shape[1])


def _zeros_like_batch(x):
    n = x.shape[0]
    if isinstance(x, torch.Tensor):
        return torch.zeros(n, dtype=torch.float32, device=x.device)
    return np.zeros(n, np.float32)


class Bijector(Module, ABC):
    """Bijector base class (bijectors.py:28-87)."""

    @abstractmethod
    def __call__(self, x, c=None, train: bool = False):
        """x (N, D), c (N, K) or None -> (y (N, D), log_det (N,))."""
        raise NotImplementedError

    @abstractmethod
    def inverse(self, x, c=None):
        """Base -> target samples; the log-determinant is not returned (bijectors.py:71)."""
        raise NotImplementedError

    # ---- native plumbing ------------------------------------------------------------------
    def _emit(self, spec: ChainSpec, scope: Scope) -> None:
        """Append this bijector's eval-mode ops to ``spec`` (reads variables from ``scope``)."""
        raise NotImplementedError(f"{type(self).__name__} has no native implementation")

    def _init_variables(self, scope: Scope, dim: int, cdim: int) -> None:
        """Create this bijector's variables (called while initializing)."""

    def _eval_forward(self, x, c):
        D = _shape2(x)[1]
        spec = ChainSpec(D, _cdim(c))
        self._emit(spec, self.scope)
        y, ld = spec.forward(x, c)
        return like_input(y, x), like_input(ld, x)

    def _eval_inverse(self, z, c):
        D = _shape2(z)[1]
        spec = ChainSpec(D, _cdim(c))
        self._emit(spec, self.scope)
        return like_input(spec.inverse(z, c), z)


class Chain(Bijector, _SequenceABC):
    """Chain of other bijectors (bijectors.py:90-124): forward applies them in order and sums
    the log-determinants, inverse applies the inverses in reverse order."""

    def __init__(self, bijectors: Sequence[Bijector]):
        self.bijectors = tuple(bijectors)

    def _children(self):
        for i, b in enumerate(self.bijectors):
            yield f"bijectors_{i}", b

    def _init_variables(self, scope, dim, cdim):
        for name, b in self._children():
            b._init_variables(scope.child(name), dim, cdim)

    def _emit(self, spec, scope):
        for name, b in self._children():
            b._emit(spec, scope.child(name))

    def __call__(self, x, c=None, train: bool = False):
        D = _shape2(x)[1]
        if self.is_initializing():
            self._init_variables(self.scope, D, _cdim(c))
            return x, _zeros_like_batch(x)
        if not train:
            return self._eval_forward(x, c)
        log_det = None
        for name, b in self._children():
            with b._bound(self.scope.child(name)):
                x, ld = b(x, c, train)
            log_det = ld if log_det is None else log_det + ld
        if log_det is None:
            log_det = _zeros_like_batch(x)
        return x, log_det

    def inverse(self, x, c=None):
        return self._eval_inverse(x, c)

    def __getitem__(self, idx: Union[int, slice]):
        """Get bijector at location idx."""
        return self.bijectors[idx]

    def __len__(self):
        """Return number of bijectors in the chain."""
        return len(self.bijectors)


def chain(*bijectors):
    """Create a chain directly from a variable number of bijector arguments (bijectors.py:127-129)."""
    return Chain(bijectors)


def _is_set(x: Optional[float]) -> bool:
    """bijectors.py:426-427."""
    return x is not None and bool(np.isfinite(x))


class ShiftBounds(Bijector):
    """Shift values into the unit interval (bijectors.py:132-273).

    Tracks the smallest and largest inputs per column in ``batch_stats/xmin_i, xmax_i`` and maps
    affinely into the unit hypercube; declared two-sided bounds are used as they are, one-sided
    bounds go through a log transform first.
    """

    def __init__(self, margin: float = 0.1,
                 bounds: Sequence[Tuple[int, Optional[float], Optional[float]]] = ()):
        self.margin = margin
        self.bounds = tuple(tuple(b) for b in bounds)

    def setup(self):
        """bijectors.py:155-161 (runs when the module is bound, as FLAX's setup does)."""
        if self.margin < 0:
            raise ValueError(f"margin must be positive (margin={self.margin})")
        if self.margin >= 1.0:
            raise ValueError(f"margin must be less than 1 (margin={self.margin})")

    def _column_kinds(self, D: int):
        bmap = {i: (a, b) for (i, a, b) in self.bounds}
        kinds, lo, hi = [], [], []
        for i in range(D):
            a, b = bmap.get(i, (None, None))
            if _is_set(a) and _is_set(b):
                k = _lib.BOUND_BOTH
            elif _is_set(a):
                k = _lib.BOUND_LOWER
            elif _is_set(b):
                k = _lib.BOUND_UPPER
            else:
                k = _lib.BOUND_NONE
            kinds.append(k)
            lo.append(float(a) if _is_set(a) else 0.0)
            hi.append(float(b) if _is_set(b) else 0.0)
        return kinds, lo, hi

    def _init_variables(self, scope, dim, cdim):
        self.setup()
        for i, a, b in self.bounds:  # bijectors.py:167-174
            if i >= dim:
                raise ValueError(f"index {i} is out of bounds")
            if _is_set(a) and _is_set(b) and b < a:
                raise ValueError("upper bound must be larger than lower bound")
        kinds, _, _ = self._column_kinds(dim)
        for i in range(dim):
            if kinds[i] != _lib.BOUND_BOTH:  # bijectors.py:243-248
                scope.variable("batch_stats", f"xmin_{i}", lambda: np.full((1,), np.inf, np.float32))
                scope.variable("batch_stats", f"xmax_{i}", lambda: np.full((1,), -np.inf, np.float32))

    def _packed_stats(self, spec: ChainSpec, scope: Scope, kinds):
        """Pack the (1,)-shaped xmin_i / xmax_i leaves into two (D,) device arrays."""
        mins, maxs = [], []
        for i, k in enumerate(kinds):
            if k == _lib.BOUND_BOTH:
                mins.append(None)
                maxs.append(None)
            else:
                mins.append(scope.get("batch_stats", f"xmin_{i}"))
                maxs.append(scope.get("batch_stats", f"xmax_{i}"))

        def pack(vals):
            if all(v is None or not isinstance(v, torch.Tensor) for v in vals):
                host = np.array([0.0 if v is None else float(np.asarray(v).reshape(-1)[0]) for v in vals], np.float32)
                return spec.leaf(host)
            parts = [torch.zeros(1, device=spec.device) if v is None else to_device_f32(v, spec.device).reshape(1)
                     for v in vals]
            return spec.leaf(torch.cat(parts))

        return pack(mins), pack(maxs)

    def _emit(self, spec, scope):
        self.setup()
        kinds, lo, hi = self._column_kinds(spec.dim)
        xmin, xmax = self._packed_stats(spec, scope, kinds)
        spec.add_shift_bounds(kinds, lo, hi, self.margin, xmin, xmax)

    def __call__(self, x, c=None, train: bool = False):
        D = _shape2(x)[1]
        if self.is_initializing():
            self._init_variables(self.scope, D, _cdim(c))
            return x, _zeros_like_batch(x)
        if not train:
            return self._eval_forward(x, c)
        from . import _train  # train-mode kernels (batch min/max), bijectors.py:250-260

        return _train.shift_bounds_train(self, x)

    def inverse(self, z, c=None):
        return self._eval_inverse(z, c)


class Roll(Bijector):
    """Roll inputs along their last axis (bijectors.py:276-297)."""

    def __init__(self, shift: int = 1):
        self.shift = shift

    def _emit(self, spec, scope):
        spec.add_roll(self.shift)

    def __call__(self, x, c=None, train: bool = False):
        if self.is_initializing():
            return x, _zeros_like_batch(x)
        return self._eval_forward(x, c)

    def inverse(self, x, c=None):
        return self._eval_inverse(x, c)


def _activation(name: str, doc: str):
    def marker(x):
        raise RuntimeError(f"{name} is evaluated inside the CUDA kernels")
    marker.__name__ = marker.__qualname__ = name
    marker.__doc__ = doc
    return marker


# Markers for NeuralSplineCoupling.act (bijectors.py:319): the kernels implement the jax.nn function of the same
# name with its default arguments.  swish (the reference default) runs on the tensor-core kernels, the others on
# the fp32 FFMA kernels.
swish = _activation("swish", "jax.nn.swish = x * sigmoid(x) (the default)")
silu = swish
relu = _activation("relu", "jax.nn.relu")
tanh = _activation("tanh", "jax.nn.tanh")
sigmoid = _activation("sigmoid", "jax.nn.sigmoid")
gelu = _activation("gelu", "jax.nn.gelu(approximate=True), jax's default")
elu = _activation("elu", "jax.nn.elu(alpha=1)")
softplus = _activation("softplus", "jax.nn.softplus")
leaky_relu = _activation("leaky_relu", "jax.nn.leaky_relu(negative_slope=0.01)")


def _act_kind(act) -> int:
    """zf_act_kind of a marker above, of its name, or of a function named like one (jax.nn.relu, ...)."""
    name = act if isinstance(act, str) else getattr(act, "__name__", "")
    name = "swish" if name == "silu" else name
    if name not in _lib.ACT_KINDS:
        raise NotImplementedError(f"act={act!r}: the kernels implement {sorted(_lib.ACT_KINDS)} "
                                  "(jax.nn definitions, default arguments)")
    return _lib.ACT_KINDS[name]


class NeuralSplineCoupling(Bijector):
    """Coupling layer with rational quadratic splines (bijectors.py:300-371).

    The upper columns of x and the conditions c drive a BatchNorm -> Dense/swish MLP whose
    output parametrises one K-knot spline per lower column; values outside [0, 1] pass
    unchanged.
    """

    def __init__(self, knots: int = 16, layers: Sequence[int] = (128, 128), act=swish):
        self.knots = knots
        self.layers = tuple(layers)
        self.act = act
        self._act_kind = _act_kind(act)

    @staticmethod
    def _split(x):
        """bijectors.py:321-327."""
        x_dim = x.shape[1]
        x_split = x_dim // 2
        assert x_split > 0 and x_split < x_dim
        return x[:, :x_split], x[:, x_split:]

    def _widths(self, dim: int):
        d = dim // 2
        return list(self.layers) + [d * (3 * self.knots - 1)]

    def _init_variables(self, scope, dim, cdim):
        d = dim // 2
        assert d > 0 and d < dim
        F = dim - d + cdim
        bn = scope.child("BatchNorm_0")
        bn.variable("params", "scale", lambda: np.ones(F, np.float32))
        bn.variable("params", "bias", lambda: np.zeros(F, np.float32))
        bn.variable("batch_stats", "mean", lambda: np.zeros(F, np.float32))
        bn.variable("batch_stats", "var", lambda: np.ones(F, np.float32))
        fan_in = F
        rng = scope.rng
        for j, w in enumerate(self._widths(dim)):
            ds = scope.child(f"Dense_{j}")

            def lecun_normal(fi=fan_in, fo=w):
                # flax default kernel_init: truncated normal (|z| <= 2) with variance 1/fan_in
                std = math.sqrt(1.0 / fi) / 0.87962566103423978
                z = rng.standard_normal((fi, fo))
                bad = np.abs(z) > 2
                while bad.any():
                    z[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(z) > 2
                return (z * std).astype(np.float32)

            ds.variable("params", "kernel", lecun_normal)
            ds.variable("params", "bias", lambda fo=w: np.zeros(fo, np.float32))
            fan_in = w

    def _leaves(self, scope: Scope, dim: int):
        bn = scope.child("BatchNorm_0")
        n = len(self.layers) + 1
        kernels = [scope.child(f"Dense_{j}").get("params", "kernel") for j in range(n)]
        biases = [scope.child(f"Dense_{j}").get("params", "bias") for j in range(n)]
        return (bn.get("params", "scale"), bn.get("params", "bias"), bn.get("batch_stats", "mean"),
                bn.get("batch_stats", "var"), kernels, biases)

    def _emit(self, spec, scope):
        assert spec.dim // 2 > 0 and spec.dim // 2 < spec.dim  # bijectors.py:326
        scale, bias, mean, var, kernels, biases = self._leaves(scope, spec.dim)
        spec.add_coupling(self.knots, self.layers, scale, bias, mean, var, kernels, biases, act=self._act_kind)

    def __call__(self, x, c=None, train: bool = False):
        D = _shape2(x)[1]
        if self.is_initializing():
            self._init_variables(self.scope, D, _cdim(c))
            return x, _zeros_like_batch(x)
        if not train:
            return self._eval_forward(x, c)
        from . import _train  # train-mode kernels (batch moments), bijectors.py:342

        return _train.coupling_train_forward(self, x, c)

    def inverse(self, y, c=None):
        return self._eval_inverse(y, c)


def rolling_spline_coupling(
    dim: int,
    knots: int = 16,
    layers: Sequence[int] = (128, 128),
    margin: Optional[float] = None,
    bounds: Sequence[Tuple[int, Optional[float], Optional[float]]] = (),
    preprocessing: Optional[Sequence[Bijector]] = None,
) -> Chain:
    """Create a chain of rolling spline couplings (bijectors.py:374-423): ShiftBounds (or the
    given preprocessing), then ``dim - 1`` x (NeuralSplineCoupling, Roll) and a final coupling."""
    if dim < 2:
        raise ValueError("dim must be at least 2")
    if preprocessing is not None:
        bijectors = list(preprocessing)
    else:
        kwargs: Dict[str, object] = {}
        if margin is not None:
            kwargs["margin"] = margin
        if bounds is not None:
            kwargs["bounds"] = bounds
        bijectors = [ShiftBounds(**kwargs)]
    for _ in range(dim - 1):
        bijectors.append(NeuralSplineCoupling(knots=knots, layers=layers))
        bijectors.append(Roll())
    bijectors.append(NeuralSplineCoupling(knots=knots, layers=layers))
    # the last Roll is skipped: the latent distribution is invariant to it
    return Chain(bijectors)
