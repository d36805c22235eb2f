"""Gradient oracle for the train step.  TEST INFRASTRUCTURE ONLY (see zenflow_oracle.py).

The reference differentiates its forward pass with ``jax.grad`` (train.py:80-86); it holds no
hand-written backward.  This file restates the *train-mode* forward (batch-statistics
BatchNorm, batch min/max ShiftBounds, ``loss = -mean(log_prob)``, train.py:64-73) with torch
float64 ops so that torch autograd plays the part of ``jax.grad``.  It is validated against
the numpy oracle (tests/test_oracle_kat.py::test_torch_oracle_matches_numpy) and is used only
as the expected value of the CUDA backward kernels.  Also here: the optax ``nadamw`` /
``adamw`` update restated from optax's documented formula (parity unpinned: optax is an
unpinned third-party dependency, pyproject.toml:11, and no reference test fixes its numbers).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

EPS = 1e-5
BN_MOMENTUM, BN_EPS = 0.99, 1e-5


def _squareplus(x):
    return 0.5 * (x + torch.sqrt(x * x + 4))


def _softmax_thr(x, thr=EPS):
    x = _squareplus(x)
    n = x.shape[-1]
    c = thr / (1 - n * thr)
    return (x / x.sum(-1, keepdim=True) + c) / (1 + c * n)


def _knots(dx):
    return torch.nn.functional.pad(torch.cumsum(dx, -1), (1, 0))


def rqs_forward(x, theta, K):
    """utils.py:37-141 on raw theta (M, d, 3K-1); in-range bins only (idx <= K-1)."""
    dx = _softmax_thr(theta[..., :K])
    dy = _softmax_thr(theta[..., K:2 * K])
    sl = _squareplus(theta[..., 2 * K:])
    xk, yk = _knots(dx), _knots(dy)
    one = torch.ones_like(sl[..., :1])
    dk = torch.cat([one, sl, one], -1)
    sk = dy / dx
    oob = (x < 0) | (x >= 1)
    idx = ((xk <= x[..., None]).sum(-1, keepdim=True) - 1).clamp(0, K - 1)
    g = lambda a, i: torch.take_along_dim(a, i, -1)[..., 0]
    xkk, ykk, dxk, dyk, dkk, dkp1, skk = g(xk, idx), g(yk, idx), g(dx, idx), g(dy, idx), g(dk, idx), g(dk, idx + 1), g(sk, idx)
    z = ((x - xkk) / dxk).clamp(EPS, 1 - EPS)
    az = 1 - z
    num = dyk * z * (skk * z + dkk * az)
    den = skk + (dkp1 + dkk - 2 * skk) * z * az
    y = ykk + num / (den + EPS)
    y = torch.where(oob, x, y)
    num2 = z * (dkp1 * z + 2 * skk * az) + dkk * az ** 2
    ld = 2 * torch.log(skk + EPS) + torch.log(num2 + EPS) - 2 * torch.log(den + EPS)
    ld = torch.where(oob, torch.zeros_like(ld), ld)
    return y, ld.sum(1)


def _batchnorm_train(h, scale, bias):
    mean = h.mean(0)
    var = torch.clamp((h * h).mean(0) - mean * mean, min=0)
    return (h - mean) * (torch.rsqrt(var + BN_EPS) * scale) + bias, mean, var


def _batchnorm_eval(h, scale, bias, mean, var):
    return (h - mean) * (torch.rsqrt(var + BN_EPS) * scale) + bias


def _shift_bounds_train(x, op, stats):
    """bijectors.py:164-273 in train mode (batch min/max merged with the running values)."""
    bmap = {i: (a, b) for (i, a, b) in op["bounds"]}
    isset = lambda v: v is not None and np.isfinite(v)
    tiny = torch.finfo(torch.float32).smallest_normal
    cols, ld = [], torch.zeros(x.shape[0], dtype=x.dtype)
    new_stats = {}
    for i in range(x.shape[1]):
        xi = x[:, i]
        a, b = bmap.get(i, (None, None))
        if isset(a) and isset(b):
            mul = 1 / (b - a)
            cols.append((xi - a) * mul)
            ld = ld + math.log(mul)
            continue
        t = xi
        if isset(a):
            t = torch.log(xi - a + tiny)
        elif isset(b):
            t = torch.log(b - xi + tiny)
        xmin, xmax = t.min(), t.max()
        delta = 0.5 * (xmax - xmin) * op["margin"]
        xmin, xmax = xmin - delta, xmax + delta
        xmin = torch.minimum(torch.as_tensor(float(stats[f"xmin_{i}"][0]), dtype=x.dtype), xmin)
        xmax = torch.maximum(torch.as_tensor(float(stats[f"xmax_{i}"][0]), dtype=x.dtype), xmax)
        new_stats[f"xmin_{i}"] = xmin.detach().numpy().reshape(1)
        new_stats[f"xmax_{i}"] = xmax.detach().numpy().reshape(1)
        mul = 1 / (xmax - xmin)
        cols.append(((t - xmin) * mul).clamp(0, 1))
        l = torch.log(mul)
        ld = ld + (l - t if (isset(a) or isset(b)) else l)
    return torch.stack(cols, 1), ld, new_stats


def _latent(z, kind, peakness):
    if kind == "beta":
        lp = (peakness - 1) * torch.log(z) + (peakness - 1) * torch.log1p(-z) - (
            2 * math.lgamma(peakness) - math.lgamma(2 * peakness))
        lp = torch.where((z > 1) | (z < 0), torch.full_like(lp, -math.inf), lp)
    elif kind in ("normal", "truncnorm"):
        lp = -(math.log(2 * math.pi * 0.01) + (z - 0.5) ** 2 / 0.01) / 2
        if kind == "truncnorm":
            lp = lp - math.log(0.5 * (math.erf(5 / math.sqrt(2)) - math.erf(-5 / math.sqrt(2))))
            lp = torch.where((z > 1) | (z < 0), torch.full_like(lp, -math.inf), lp)
    else:
        lp = torch.where((z > 1) | (z < 0), torch.full_like(z, -math.inf), torch.zeros_like(z))
    return lp.sum(-1)


def to_torch(tree, requires_grad=False):
    if isinstance(tree, dict):
        return {k: to_torch(v, requires_grad) for k, v in tree.items()}
    t = torch.tensor(np.asarray(tree, np.float64))
    return t.requires_grad_(requires_grad)


def to_numpy(tree):
    if isinstance(tree, dict):
        return {k: to_numpy(v) for k, v in tree.items()}
    return tree.detach().numpy()


# bijectors.py:319 `act` (jax.nn definitions, default arguments)
_ACT = {
    "swish": lambda h: h * torch.sigmoid(h),
    "relu": torch.relu,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "gelu": lambda h: torch.nn.functional.gelu(h, approximate="tanh"),
    "elu": torch.nn.functional.elu,
    "softplus": lambda h: torch.logaddexp(h, torch.zeros_like(h)),   # F.softplus switches to x above 20
    "leaky_relu": lambda h: torch.nn.functional.leaky_relu(h, 0.01),
}


def train_loss(ops, params_t, stats, x, c, *, latent="beta", peakness=12.0, train=True):
    """loss_fn of train.py:64-73: returns (loss, new_batch_stats, lp).  x, c torch float64."""
    new_stats: Dict[str, dict] = {}
    ld_total = torch.zeros(x.shape[0], dtype=x.dtype)
    for i, op in enumerate(ops):
        name = f"bijectors_{i}"
        if op["kind"] == "shift_bounds":
            if train:
                x, ld, ns = _shift_bounds_train(x, op, stats[name])
                new_stats[name] = {**stats[name], **ns}
            else:
                raise NotImplementedError
            ld_total = ld_total + ld
        elif op["kind"] == "roll":
            x = torch.roll(x, op["shift"], -1)
        else:
            p, s = params_t[name], stats[name]
            K = op["knots"]
            D = x.shape[1]
            d = D // 2
            xt, xc = x[:, :d], x[:, d:]
            h = torch.cat([xc, c], 1) if c is not None else xc
            bn = p["BatchNorm_0"]
            if train:
                h, mean, var = _batchnorm_train(h, bn["scale"], bn["bias"])
                rm, rv = np.asarray(s["BatchNorm_0"]["mean"], np.float64), np.asarray(s["BatchNorm_0"]["var"], np.float64)
                new_stats[name] = {"BatchNorm_0": {
                    "mean": BN_MOMENTUM * rm + (1 - BN_MOMENTUM) * mean.detach().numpy(),
                    "var": BN_MOMENTUM * rv + (1 - BN_MOMENTUM) * var.detach().numpy()}}
            else:
                h = _batchnorm_eval(h, bn["scale"], bn["bias"], torch.tensor(np.asarray(s["BatchNorm_0"]["mean"], np.float64)),
                                    torch.tensor(np.asarray(s["BatchNorm_0"]["var"], np.float64)))
            n_dense = sum(1 for k in p if k.startswith("Dense_"))
            for j in range(n_dense - 1):
                h = h @ p[f"Dense_{j}"]["kernel"] + p[f"Dense_{j}"]["bias"]
                h = _ACT[op.get("act", "swish")](h)
            j = n_dense - 1
            theta = (h @ p[f"Dense_{j}"]["kernel"] + p[f"Dense_{j}"]["bias"]).reshape(x.shape[0], d, 3 * K - 1)
            yt, ld = rqs_forward(xt, theta, K)
            x = torch.cat([yt, xc], 1)
            ld_total = ld_total + ld
    lp = _latent(x, latent, peakness) + ld_total
    lp = torch.nan_to_num(lp, nan=-math.inf, posinf=torch.finfo(torch.float32).max, neginf=torch.finfo(torch.float32).min)
    return -lp.mean(), new_stats, lp


def loss_and_grads(ops, variables, x, c, *, latent="beta", peakness=12.0):
    """jax.grad(loss_fn, has_aux=True) of train.py:82 (+ d loss / d c for cfg3)."""
    params_t = to_torch(variables["params"], requires_grad=True)
    xt = torch.tensor(np.asarray(x, np.float64))
    ct = None if c is None else torch.tensor(np.asarray(c, np.float64)).requires_grad_(True)
    loss, new_stats, lp = train_loss(ops, params_t, variables["batch_stats"], xt, ct, latent=latent, peakness=peakness)
    loss.backward()
    grads = _grads_of(params_t)
    gc = None if ct is None else ct.grad.numpy()
    return float(loss.detach()), grads, new_stats, gc, lp.detach().numpy()


def _grads_of(tree):
    if isinstance(tree, dict):
        return {k: _grads_of(v) for k, v in tree.items()}
    return np.zeros(tuple(tree.shape)) if tree.grad is None else tree.grad.numpy()


def nadamw_update(params, grads, mu, nu, count, *, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-4,
                  nesterov=True):
    """optax.nadamw (= adamw(nesterov=True)): scale_by_adam -> add_decayed_weights -> scale(-lr).
    Works leaf-wise on numpy arrays; returns (new_params, new_mu, new_nu, count+1)."""
    t = count + 1
    mu2 = b1 * mu + (1 - b1) * grads
    nu2 = b2 * nu + (1 - b2) * grads * grads
    if nesterov:
        mu_hat = b1 * (mu2 / (1 - b1 ** (t + 1))) + (1 - b1) * (grads / (1 - b1 ** t))
    else:
        mu_hat = mu2 / (1 - b1 ** t)
    nu_hat = nu2 / (1 - b2 ** t)
    upd = mu_hat / (np.sqrt(nu_hat) + eps) + weight_decay * params
    return params - lr * upd, mu2, nu2, t
