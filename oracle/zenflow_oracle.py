"""CPU oracle for the zenflow spline-coupling hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm (HDembinski/zenflow,
pure JAX/FLAX).  It is the *checker* for the CUDA path; nothing in the product
package ``zenflow_b200`` may import it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it.

Parity status
-------------
* pinned against every known-answer test the reference holds for this path
  (``tests/test_utils.py``, ``tests/test_bijectors.py``, ``tests/test_distributions.py``
  of the reference; see ``tests/test_oracle_kat.py`` here);
* **parity unpinned against a running JAX**: jax/jaxlib/flax/optax are not installed
  in this image and there is no network, so the reference itself cannot be executed
  (SURVEY.md F2).  Third-party semantics (flax ``Dense``/``BatchNorm``, ``jax.nn.swish``,
  ``jax.scipy.stats`` log-pdfs, ``jnp.nan_to_num``, ``jnp.take_along_axis`` fill mode,
  XLA's reduction order) are restated from their documented behaviour and marked
  ``UNVERIFIED`` where the reference's tests do not pin them.

Conventions
-----------
* every function takes/returns numpy arrays and computes in the dtype of its
  floating inputs (float32 = the reference's arithmetic, float64 = truth mode);
* reductions over the knot axis (``jnp.sum`` / ``jnp.cumsum`` in ``utils.py:33,237``)
  are **sequential left-to-right** in the working dtype (UNVERIFIED vs XLA); the CUDA
  kernels use the same order so that bin indices are bit-exact given identical
  raw spline parameters;
* no fused multiply-add anywhere on the knot path (numpy never contracts).

All ``file:line`` citations are relative to ``/root/reference/src/zenflow``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

EPS = 1e-5  # utils.py:15

__all__ = [
    "EPS",
    "squareplus",
    "softmax_with_threshold",
    "normalize_spline_params",
    "knots",
    "index",
    "compute_rqs_input",
    "rqs_forward",
    "rqs_inverse",
    "split_theta",
    "rqs_forward_theta",
    "rqs_inverse_theta",
    "swish",
    "dense",
    "batchnorm",
    "roll_forward",
    "roll_inverse",
    "shift_bounds_forward",
    "shift_bounds_inverse",
    "coupling_params",
    "coupling_forward",
    "coupling_inverse",
    "chain_forward",
    "chain_inverse",
    "latent_log_prob",
    "nan_to_num",
    "flow_log_prob",
    "flow_inverse",
    "init_variables",
    "make_chain",
]


# ----------------------------------------------------------------------------------
# utils.py — spline math
# ----------------------------------------------------------------------------------


def _seq_sum(x: np.ndarray) -> np.ndarray:
    """Sequential left-to-right sum over the last axis in x.dtype (keeps the axis)."""
    return np.add.accumulate(x, axis=-1)[..., -1:]


def squareplus(x: np.ndarray, b: float = 4) -> np.ndarray:
    """utils.py:18-20  ``0.5 * (x + sqrt(square(x) + b))``."""
    dt = x.dtype.type
    return dt(0.5) * (x + np.sqrt(x * x + dt(b)))


def softmax_with_threshold(x: np.ndarray, threshold: float = 0) -> np.ndarray:
    """utils.py:23-34.  ``c`` and ``1 + c*n`` are Python doubles cast to x.dtype."""
    dt = x.dtype.type
    x = squareplus(x)
    n = x.shape[-1]
    c = threshold / (1 - n * threshold)
    xs = _seq_sum(x)
    return (x / xs + dt(c)) / dt(1 + c * n)


def normalize_spline_params(dx, dy, sl):
    """utils.py:37-62."""
    return (
        softmax_with_threshold(dx, EPS),
        softmax_with_threshold(dy, EPS),
        squareplus(sl),
    )


def knots(dx: np.ndarray) -> np.ndarray:
    """utils.py:235-241  pad(cumsum(dx), left 0): K+1 knot positions."""
    cs = np.add.accumulate(dx, axis=-1)
    pad = np.zeros(dx.shape[:-1] + (1,), dtype=dx.dtype)
    return np.concatenate([pad, cs], axis=-1)


def index(x: np.ndarray, xk: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:244-250.  Returns idx (..., 1) int32 in [0, K] and the oob mask."""
    with np.errstate(invalid="ignore"):
        oob = (x < 0) | (x >= 1)
        idx = np.sum(xk <= x[..., None], axis=-1, dtype=np.int32)[..., None] - 1
    idx = np.clip(idx, 0, xk.shape[-1] - 1).astype(np.int32)
    return idx, oob


def _take_fill(arr: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """``jnp.take_along_axis(arr, idx, -1)[..., 0]`` with JAX's default
    out-of-bounds mode "fill" (NaN for floats) — UNVERIFIED."""
    n = arr.shape[-1]
    bad = (idx < 0) | (idx >= n)
    safe = np.where(bad, 0, idx)
    out = np.take_along_axis(arr, safe, axis=-1)
    out = np.where(bad, arr.dtype.type(np.nan), out)
    return out[..., 0]


def compute_rqs_input(x, dx, dy, slope, forward: bool):
    """utils.py:205-232."""
    xk = knots(dx)
    yk = knots(dy)
    one = np.ones(slope.shape[:-1] + (1,), dtype=slope.dtype)
    dk = np.concatenate([one, slope, one], axis=-1)  # :211-216 boundary derivative 1
    sk = dy / dx  # :218
    idx, oob = index(x, xk if forward else yk)
    return (
        _take_fill(xk, idx),
        _take_fill(yk, idx),
        _take_fill(dx, idx),
        _take_fill(dy, idx),
        _take_fill(dk, idx),
        _take_fill(dk, idx + 1),
        _take_fill(sk, idx),
        oob,
        idx[..., 0],
    )


def rqs_forward(x, dx, dy, slope, return_idx: bool = False):
    """utils.py:65-141.  x (M,N); dx,dy (M,N,K); slope (M,N,K-1) -> y (M,N), log_det (M,)."""
    dt = x.dtype.type
    xk, yk, dxk, dyk, dk, dkp1, sk, oob, idx = compute_rqs_input(x, dx, dy, slope, True)
    eps = dt(EPS)
    with np.errstate(all="ignore"):
        z = (x - xk) / dxk
        z = np.clip(z, eps, dt(1 - EPS))
        az = dt(1) - z
        num = dyk * z * (sk * z + dk * az)
        den = sk + (dkp1 + dk - dt(2) * sk) * z * az
        y = yk + num / (den + eps)
        y = np.where(oob, x, y)
        num = z * (dkp1 * z + dt(2) * sk * az) + dk * (az * az)
        den = sk + (dkp1 + dk - dt(2) * sk) * z * az
        log_det = dt(2) * np.log(sk + eps) + np.log(num + eps) - dt(2) * np.log(den + eps)
        log_det = np.where(oob, dt(0), log_det)
    log_det = log_det.sum(axis=1, dtype=x.dtype)
    if return_idx:
        return y, log_det, idx
    return y, log_det


def rqs_inverse(y, dx, dy, slope, return_idx: bool = False):
    """utils.py:144-202.  No EPS, no z-clip, no log-det."""
    dt = y.dtype.type
    xk, yk, dxk, dyk, dk, dkp1, sk, oob, idx = compute_rqs_input(y, dx, dy, slope, False)
    with np.errstate(all="ignore"):
        beta = dkp1 + dk - dt(2) * sk
        a = dyk * (sk - dk) + (y - yk) * beta
        b = dyk * dk - (y - yk) * beta
        c = -sk * (y - yk)
        z = dt(2) * c / (-b - np.sqrt(b * b - dt(4) * a * c))
        x = z * dxk + xk
        x = np.where(oob, y, x)
    if return_idx:
        return x, idx
    return x


def split_theta(theta: np.ndarray, K: int):
    """bijectors.py:353-355: raw conditioner output (M,d,3K-1) -> widths, heights, slopes."""
    return theta[..., :K], theta[..., K : 2 * K], theta[..., 2 * K :]


def rqs_forward_theta(x, theta, K: int, return_idx: bool = False):
    """normalize_spline_params + rational_quadratic_spline_forward on raw theta."""
    dx, dy, sl = normalize_spline_params(*split_theta(theta, K))
    return rqs_forward(x, dx, dy, sl, return_idx)


def rqs_inverse_theta(y, theta, K: int, return_idx: bool = False):
    dx, dy, sl = normalize_spline_params(*split_theta(theta, K))
    return rqs_inverse(y, dx, dy, sl, return_idx)


# ----------------------------------------------------------------------------------
# third-party layer semantics (flax.linen / jax.nn) — restated, UNVERIFIED
# ----------------------------------------------------------------------------------


def swish(x: np.ndarray) -> np.ndarray:
    """jax.nn.swish = x * sigmoid(x), sigmoid = 1/(1+exp(-x))."""
    dt = x.dtype.type
    with np.errstate(over="ignore"):
        return x * (dt(1) / (dt(1) + np.exp(-x)))


def dense(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """flax.linen.Dense: x @ kernel + bias, kernel (in, out)."""
    return x @ kernel.astype(x.dtype) + bias.astype(x.dtype)


BN_MOMENTUM = 0.99  # flax.linen.BatchNorm defaults
BN_EPS = 1e-5


def batchnorm(x, scale, bias, ra_mean, ra_var, train: bool):
    """flax.linen.BatchNorm(use_running_average=not train), feature axis -1.

    Train: batch mean and biased variance computed as E[x^2]-E[x]^2 clipped at 0
    (use_fast_variance=True), running stats updated with momentum 0.99.
    Returns (y, new_mean, new_var).
    """
    dt = x.dtype.type
    if train:
        mean = x.mean(axis=0, dtype=x.dtype)
        mean2 = (x * x).mean(axis=0, dtype=x.dtype)
        var = np.maximum(dt(0), mean2 - mean * mean)
        new_mean = dt(BN_MOMENTUM) * ra_mean.astype(x.dtype) + dt(1 - BN_MOMENTUM) * mean
        new_var = dt(BN_MOMENTUM) * ra_var.astype(x.dtype) + dt(1 - BN_MOMENTUM) * var
    else:
        mean = ra_mean.astype(x.dtype)
        var = ra_var.astype(x.dtype)
        new_mean, new_var = ra_mean, ra_var
    mul = (dt(1) / np.sqrt(var + dt(BN_EPS))) * scale.astype(x.dtype)
    y = (x - mean) * mul + bias.astype(x.dtype)
    return y, new_mean, new_var


# ----------------------------------------------------------------------------------
# bijectors.py
# ----------------------------------------------------------------------------------


def roll_forward(x: np.ndarray, shift: int = 1) -> np.ndarray:
    """bijectors.py:288-293  jnp.roll(x, shift, axis=-1)."""
    return np.roll(x, shift, axis=-1)


def roll_inverse(x: np.ndarray, shift: int = 1) -> np.ndarray:
    """bijectors.py:295-297."""
    return np.roll(x, -shift, axis=-1)


def _is_set(v) -> bool:
    """bijectors.py:426-427."""
    return v is not None and np.isfinite(v)


def _safe_log(x: np.ndarray) -> np.ndarray:
    """bijectors.py:430-431."""
    with np.errstate(all="ignore"):
        return np.log(x + np.finfo(x.dtype).smallest_normal)


def _unit_interval(stats: Dict[str, np.ndarray], i: int, x: np.ndarray, train: bool,
                   margin: float, initializing: bool):
    """bijectors.py:242-273.  stats holds xmin_i / xmax_i arrays of shape (1,)."""
    dt = x.dtype.type
    ra_min = stats.setdefault(f"xmin_{i}", np.full((1,), np.inf, dtype=np.float32))
    ra_max = stats.setdefault(f"xmax_{i}", np.full((1,), -np.inf, dtype=np.float32))
    with np.errstate(all="ignore"):
        if train:
            xmin = x.min()
            xmax = x.max()
            xdelta = dt(0.5) * (xmax - xmin) * dt(margin)
            xmin = xmin - xdelta
            xmax = xmax + xdelta
            xmin = np.minimum(ra_min.astype(x.dtype), xmin)
            xmax = np.maximum(ra_max.astype(x.dtype), xmax)
            if not initializing:
                stats[f"xmin_{i}"] = xmin.astype(np.float32).reshape(1)
                stats[f"xmax_{i}"] = xmax.astype(np.float32).reshape(1)
        else:
            xmin = ra_min.astype(x.dtype)
            xmax = ra_max.astype(x.dtype)
        mul = dt(1) / (xmax - xmin)
        z = (x - xmin) * mul
        ld = np.log(mul)
        z = np.clip(z, dt(0), dt(1))
    return z, ld


def shift_bounds_forward(x, stats: Dict[str, np.ndarray], *, margin: float = 0.1,
                         bounds: Sequence[Tuple[int, Optional[float], Optional[float]]] = (),
                         train: bool = False, initializing: bool = False):
    """bijectors.py:164-208.  ``stats`` is updated in place when train and not initializing.
    Returns (z, log_det)."""
    if x.dtype.kind == "i":
        x = x.astype(np.float32)
    dt = x.dtype.type
    bmap = {i: (a, b) for (i, a, b) in bounds}
    z = np.empty_like(x)
    log_det = np.zeros(x.shape[0], x.dtype)
    for i in range(x.shape[1]):
        xi = x[:, i]
        a, b = bmap.get(i, (None, None))
        if _is_set(a):
            if _is_set(b):
                mul = 1 / (b - a)
                zi = (xi - dt(a)) * dt(mul)
                ld = np.log(dt(mul))
            else:
                ti = _safe_log(xi - dt(a))
                zi, ld = _unit_interval(stats, i, ti, train, margin, initializing)
                ld = ld - ti
        elif _is_set(b):
            ti = _safe_log(dt(b) - xi)
            zi, ld = _unit_interval(stats, i, ti, train, margin, initializing)
            ld = ld - ti
        else:
            zi, ld = _unit_interval(stats, i, xi, train, margin, initializing)
        z[:, i] = zi
        log_det = log_det + ld
    return z, log_det


def shift_bounds_inverse(z, stats: Dict[str, np.ndarray], *,
                         bounds: Sequence[Tuple[int, Optional[float], Optional[float]]] = ()):
    """bijectors.py:210-240."""
    dt = z.dtype.type
    bmap = {i: (a, b) for (i, a, b) in bounds}
    x = np.empty_like(z)
    with np.errstate(all="ignore"):
        for i in range(z.shape[1]):
            zi = z[:, i]
            a, b = bmap.get(i, (None, None))
            if _is_set(a) and _is_set(b):
                xi = zi * dt(b) + (dt(1) - zi) * dt(a)
            else:
                xmin = stats[f"xmin_{i}"].astype(z.dtype)
                xmax = stats[f"xmax_{i}"].astype(z.dtype)
                ti = zi * xmax + (dt(1) - zi) * xmin
                if _is_set(a):
                    xi = np.exp(ti) + dt(a)
                elif _is_set(b):
                    xi = dt(b) - np.exp(ti)
                else:
                    xi = ti
            x[:, i] = xi
    return x


def relu(x):
    """jax.nn.relu."""
    return np.maximum(x, x.dtype.type(0))


def sigmoid(x):
    """jax.nn.sigmoid."""
    dt = x.dtype.type
    with np.errstate(over="ignore"):
        return dt(1) / (dt(1) + np.exp(-x))


def gelu(x):
    """jax.nn.gelu(approximate=True) (jax's default): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))."""
    dt = x.dtype.type
    return dt(0.5) * x * (dt(1) + np.tanh(dt(math.sqrt(2.0 / math.pi)) * (x + dt(0.044715) * (x * x * x))))


def elu(x):
    """jax.nn.elu(alpha=1)."""
    with np.errstate(over="ignore"):
        return np.where(x > 0, x, np.expm1(x)).astype(x.dtype)


def softplus(x):
    """jax.nn.softplus = logaddexp(x, 0)."""
    return np.logaddexp(x, x.dtype.type(0))


def leaky_relu(x):
    """jax.nn.leaky_relu(negative_slope=0.01)."""
    return np.where(x >= 0, x, x.dtype.type(0.01) * x).astype(x.dtype)


# bijectors.py:319 `act` by the name of the jax.nn function
ACTIVATIONS = {"swish": swish, "silu": swish, "relu": relu, "tanh": np.tanh, "sigmoid": sigmoid, "gelu": gelu,
               "elu": elu, "softplus": softplus, "leaky_relu": leaky_relu}


def coupling_params(x, c, params, stats, *, knots_: int, train: bool, act: str = "swish"):
    """bijectors.py:329-357  the conditioner.  Returns (xt, xc, theta, new_stats).

    params: {"BatchNorm_0": {"scale","bias"}, "Dense_j": {"kernel","bias"}}
    stats:  {"BatchNorm_0": {"mean","var"}}
    theta is the raw (M, d, 3K-1) output before normalize_spline_params.
    """
    D = x.shape[1]
    d = D // 2
    assert 0 < d < D  # :326
    xt, xc = x[:, :d], x[:, d:]
    h = np.hstack((xc, c.astype(x.dtype))) if c is not None else xc
    bn_p, bn_s = params["BatchNorm_0"], stats["BatchNorm_0"]
    h, m, v = batchnorm(h, bn_p["scale"], bn_p["bias"], bn_s["mean"], bn_s["var"], train)
    n_dense = sum(1 for k in params if k.startswith("Dense_"))
    for j in range(n_dense - 1):
        h = ACTIVATIONS[act](dense(h, params[f"Dense_{j}"]["kernel"], params[f"Dense_{j}"]["bias"]))  # :345
    j = n_dense - 1
    h = dense(h, params[f"Dense_{j}"]["kernel"], params[f"Dense_{j}"]["bias"])
    theta = h.reshape(x.shape[0], d, 3 * knots_ - 1)
    new_stats = {"BatchNorm_0": {"mean": np.asarray(m, np.float32), "var": np.asarray(v, np.float32)}}
    return xt, xc, theta, new_stats


def coupling_forward(x, c, params, stats, *, knots_: int, train: bool = False,
                     return_aux: bool = False, act: str = "swish"):
    """bijectors.py:359-365."""
    xt, xc, theta, new_stats = coupling_params(x, c, params, stats, knots_=knots_, train=train, act=act)
    yt, log_det, idx = rqs_forward_theta(xt, theta, knots_, return_idx=True)
    y = np.hstack((yt, xc))
    if return_aux:
        return y, log_det, new_stats, {"theta": theta, "idx": idx}
    return y, log_det, new_stats


def coupling_inverse(y, c, params, stats, *, knots_: int, return_aux: bool = False, act: str = "swish"):
    """bijectors.py:367-371 (always eval-mode BatchNorm)."""
    yt, yc, theta, _ = coupling_params(y, c, params, stats, knots_=knots_, train=False, act=act)
    xt, idx = rqs_inverse_theta(yt, theta, knots_, return_idx=True)
    x = np.hstack((xt, yc))
    if return_aux:
        return x, {"theta": theta, "idx": idx}
    return x


# A chain is described by a list of dicts, the oracle's stand-in for the FLAX modules:
#   {"kind": "shift_bounds", "margin": 0.1, "bounds": ()}
#   {"kind": "roll", "shift": 1}
#   {"kind": "coupling", "knots": 16, "layers": (128, 128)}      (+ "act": a key of ACTIVATIONS; default "swish")
# and variables use the FLAX naming: params["bijectors_i"], batch_stats["bijectors_i"].


def make_chain(dim: int, knots_: int = 16, layers: Sequence[int] = (128, 128),
               margin: Optional[float] = None, bounds=(), n_couplings: Optional[int] = None,
               roll_shift: int = 1) -> List[dict]:
    """bijectors.py:374-423 rolling_spline_coupling (n_couplings=None ⇒ dim couplings)."""
    if dim < 2:
        raise ValueError("dim must be at least 2")
    sb = {"kind": "shift_bounds", "margin": 0.1 if margin is None else margin,
          "bounds": tuple(bounds)}
    ops: List[dict] = [sb]
    n = dim if n_couplings is None else n_couplings
    for _ in range(n - 1):
        ops.append({"kind": "coupling", "knots": knots_, "layers": tuple(layers)})
        ops.append({"kind": "roll", "shift": roll_shift})
    ops.append({"kind": "coupling", "knots": knots_, "layers": tuple(layers)})
    return ops


def init_variables(ops: List[dict], dim: int, cdim: int, seed: int = 0, *,
                   weight_scale: float = 1.0, randomize_bn: bool = False):
    """Build a FLAX-shaped variable tree for a chain.

    Kernels: LeCun-normal-like N(0, weight_scale/fan_in) (flax default is a truncated
    normal of the same variance; the RNG stream differs from jax.random regardless),
    zero biases, BN scale 1 / bias 0 / mean 0 / var 1, ShiftBounds stats ±inf.
    ``randomize_bn`` perturbs BN params/stats and biases so that tests exercise them.
    """
    rng = np.random.default_rng(seed)
    params: Dict[str, dict] = {}
    stats: Dict[str, dict] = {}
    for i, op in enumerate(ops):
        name = f"bijectors_{i}"
        if op["kind"] == "shift_bounds":
            st = {}
            for j in range(dim):
                a, b = {k: (lo, hi) for (k, lo, hi) in op["bounds"]}.get(j, (None, None))
                if _is_set(a) and _is_set(b):
                    continue
                st[f"xmin_{j}"] = np.full((1,), np.inf, np.float32)
                st[f"xmax_{j}"] = np.full((1,), -np.inf, np.float32)
            stats[name] = st
        elif op["kind"] == "coupling":
            d = dim // 2
            F = dim - d + cdim
            widths = list(op["layers"]) + [d * (3 * op["knots"] - 1)]
            p = {"BatchNorm_0": {"scale": np.ones(F, np.float32), "bias": np.zeros(F, np.float32)}}
            s = {"BatchNorm_0": {"mean": np.zeros(F, np.float32), "var": np.ones(F, np.float32)}}
            if randomize_bn:
                p["BatchNorm_0"]["scale"] = rng.uniform(0.5, 1.5, F).astype(np.float32)
                p["BatchNorm_0"]["bias"] = rng.normal(0, 0.2, F).astype(np.float32)
                s["BatchNorm_0"]["mean"] = rng.uniform(0.3, 0.7, F).astype(np.float32)
                s["BatchNorm_0"]["var"] = rng.uniform(0.05, 0.2, F).astype(np.float32)
            fan_in = F
            for j, w in enumerate(widths):
                k = rng.normal(0.0, math.sqrt(weight_scale / fan_in), (fan_in, w)).astype(np.float32)
                b = (rng.normal(0, 0.1, w) if randomize_bn else np.zeros(w)).astype(np.float32)
                p[f"Dense_{j}"] = {"kernel": k, "bias": b}
                fan_in = w
            params[name] = p
            stats[name] = s
    return {"params": params, "batch_stats": stats}


def chain_forward(ops, variables, x, c=None, *, train: bool = False, initializing: bool = False,
                  return_steps: bool = False):
    """bijectors.py:104-111.  Returns (y, log_det, new_batch_stats[, steps])."""
    if c is not None and c.ndim == 1:
        c = c.reshape(-1, 1)  # flow.py:98-101
    params = variables.get("params", {})
    stats = {k: dict(v) for k, v in variables.get("batch_stats", {}).items()}
    log_det = np.zeros(x.shape[0], x.dtype if x.dtype.kind == "f" else np.float32)
    steps = []
    for i, op in enumerate(ops):
        name = f"bijectors_{i}"
        if op["kind"] == "shift_bounds":
            st = stats.setdefault(name, {})
            x, ld = shift_bounds_forward(x, st, margin=op["margin"], bounds=op["bounds"],
                                         train=train, initializing=initializing)
        elif op["kind"] == "roll":
            x, ld = roll_forward(x, op["shift"]), 0
        elif op["kind"] == "coupling":
            x, ld, ns = coupling_forward(x, c, params[name], stats[name], knots_=op["knots"],
                                         train=train, act=op.get("act", "swish"))
            if train and not initializing:
                stats[name] = ns
        else:
            raise ValueError(op["kind"])
        log_det = log_det + ld
        steps.append(x)
    if return_steps:
        return x, log_det, stats, steps
    return x, log_det, stats


def chain_inverse(ops, variables, z, c=None, *, return_steps: bool = False):
    """bijectors.py:113-116."""
    if c is not None and c.ndim == 1:
        c = c.reshape(-1, 1)
    params = variables.get("params", {})
    stats = variables.get("batch_stats", {})
    steps = []
    x = z
    for i in reversed(range(len(ops))):
        op = ops[i]
        name = f"bijectors_{i}"
        if op["kind"] == "shift_bounds":
            x = shift_bounds_inverse(x, stats[name], bounds=op["bounds"])
        elif op["kind"] == "roll":
            x = roll_inverse(x, op["shift"])
        elif op["kind"] == "coupling":
            x = coupling_inverse(x, c, params[name], stats[name], knots_=op["knots"], act=op.get("act", "swish"))
        steps.append(x)
    if return_steps:
        return x, steps
    return x


# ----------------------------------------------------------------------------------
# distributions.py — latent log-pdfs (jax.scipy.stats restated)
# ----------------------------------------------------------------------------------

_LOG_2PI = math.log(2 * math.pi)


def _betaln(a: float, b: float) -> float:
    return math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)


def _log_gauss_mass(a: float, b: float) -> float:
    """log(Phi(b) - Phi(a)) for the truncation interval (jax.scipy.stats.truncnorm)."""
    return math.log(0.5 * (math.erf(b / math.sqrt(2)) - math.erf(a / math.sqrt(2))))


def latent_log_prob(x: np.ndarray, kind: str = "beta", peakness: float = 12.0) -> np.ndarray:
    """distributions.py:58-59 (normal), :72-73 (truncnorm), :100-104 (beta), :122-123 (uniform).
    Sum over the last axis of the per-dimension log-pdf."""
    dt = x.dtype.type
    with np.errstate(all="ignore"):
        if kind == "beta":
            p1 = dt(peakness - 1)
            # xlogy / xlog1py: 0*log(0) = 0
            t0 = np.where((p1 == 0) & (x == 0), dt(0), p1 * np.log(x))
            t1 = np.where((p1 == 0) & (x == 1), dt(0), p1 * np.log1p(-x))
            lp = t0 + t1 - dt(_betaln(peakness, peakness))
            lp = np.where((x > 1) | (x < 0), dt(-np.inf), lp)
        elif kind in ("normal", "truncnorm"):
            scale = 0.1
            log_norm = dt(math.log(2 * math.pi * scale * scale))
            quad = (x - dt(0.5)) * (x - dt(0.5)) / dt(scale * scale)
            lp = -(log_norm + quad) / dt(2)
            if kind == "truncnorm":
                lp = lp - dt(_log_gauss_mass(-5.0, 5.0))
                lo, hi = dt(-5.0 * scale + 0.5), dt(5.0 * scale + 0.5)
                lp = np.where((x < lo) | (x > hi), dt(-np.inf), lp)
        elif kind == "uniform":
            lp = np.where((x > 1) | (x < 0), dt(-np.inf), dt(0)) + np.zeros_like(x)
        else:
            raise ValueError(kind)
    return lp.sum(axis=-1, dtype=x.dtype)


def nan_to_num(lp: np.ndarray) -> np.ndarray:
    """flow.py:47 ``jnp.nan_to_num(lp, nan=-inf)``: three sequential wheres
    (nan -> -inf, +inf -> finfo.max, -inf -> finfo.min) — UNVERIFIED ordering,
    so NaN ends at finfo.min."""
    fi = np.finfo(lp.dtype)
    out = np.where(np.isnan(lp), lp.dtype.type(-np.inf), lp)
    out = np.where(np.isposinf(out), fi.max, out)
    out = np.where(np.isneginf(out), fi.min, out)
    return out.astype(lp.dtype)


# ----------------------------------------------------------------------------------
# flow.py
# ----------------------------------------------------------------------------------


def flow_log_prob(ops, variables, x, c=None, *, latent: str = "beta", peakness: float = 12.0,
                  train: bool = False):
    """flow.py:22-48.  Returns (log_prob, new_batch_stats)."""
    z, log_det, stats = chain_forward(ops, variables, x, c, train=train)
    with np.errstate(all="ignore"):
        lp = latent_log_prob(z, latent, peakness) + log_det
    return nan_to_num(lp), stats


def flow_inverse(ops, variables, u, c=None):
    """flow.py:77: bijector.inverse(u, c) with the latent draw u given as input."""
    return chain_inverse(ops, variables, u, c)
