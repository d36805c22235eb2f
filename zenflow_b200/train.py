"""Train flow — host-side mirror of zenflow/train.py.

Same signature and bookkeeping as the reference's ``train`` (per-epoch permutation,
minibatch ``step``, best-epoch tracking, non-finite abort, patience-window early stop);
the jitted ``step`` (jax.grad of ``-mean(log_prob)`` + optax update, train.py:64-86) is the
native train step (``TrainEngine``).  The optimiser is described by a small config object
instead of an optax GradientTransformation: ``nadamw`` (the reference's default when optax
has it, train.py:12-15) or ``adamw``, same hyper-parameter names and defaults as optax.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._device import ptr, require_cuda, stream_ptr, to_device_f32
from ._train import TrainEngine, _allreduce, _dist_world
from .flow import Flow

__all__ = ["train", "nadamw", "adamw", "Optimizer", "DEFAULT_OPTIMIZER"]


@dataclass(frozen=True)
class Optimizer:
    """optax.adamw / optax.nadamw hyper-parameters (optax defaults)."""

    learning_rate: float = 1e-3
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 1e-4
    nesterov: bool = True


def nadamw(learning_rate: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
           weight_decay: float = 1e-4) -> Optimizer:
    return Optimizer(learning_rate, b1, b2, eps, weight_decay, True)


def adamw(learning_rate: float = 1e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
          weight_decay: float = 1e-4) -> Optimizer:
    return Optimizer(learning_rate, b1, b2, eps, weight_decay, False)


DEFAULT_OPTIMIZER = nadamw


def _cdim(C) -> int:
    if C is None:
        return 0
    return 1 if C.ndim == 1 else int(C.shape[1])


def train(
    flow: Flow,
    X_train,
    X_test,
    C_train=None,
    C_test=None,
    *,
    epochs: int = 1000,
    batch_size: int = 1024,
    optimizer: Optimizer = DEFAULT_OPTIMIZER(learning_rate=1e-3),
    patience: float = 0.05,
    warmup: float = 0.2,
    seed: int = 0,
    progress: bool = True,
    initial_variables=None,
    group=None,
) -> Tuple[dict, int, List[float], List[float]]:
    """Trains the normalizing flow on the provided inputs (train.py:18-138).

    Returns (best_variables, best_epoch, loss_train, loss_test); the variables are a FLAX-shaped
    pytree of device tensors.

    Data-parallel training: initialise torch.distributed (or pass ``group``) and give every rank its own row
    shard of X_train / C_train AND of X_test / C_test.  The shards must give every rank the same number of
    minibatch steps per epoch (checked up front: ``ValueError`` on every rank otherwise); the losses in the
    returned histories are global (sums and counts are all-reduced), so the non-finite abort, the best-epoch
    bookkeeping and the early stop take the same decision on every rank.
    """
    if warmup < 1:
        warmup = warmup * epochs
    warmup = int(warmup)
    if patience < 1:
        patience = patience * epochs
    patience = int(patience)

    dev = require_cuda()
    X_train = to_device_f32(X_train, dev)  # jax.device_put, train.py:43-48
    X_test = to_device_f32(X_test, dev)
    if C_train is not None:
        C_train = to_device_f32(C_train, dev)
    if C_test is not None:
        C_test = to_device_f32(C_test, dev)

    if initial_variables is None:
        variables = flow.init(seed, X_train[:1].cpu().numpy(), None if C_train is None else C_train[:1].cpu().numpy())
    else:
        variables = initial_variables
        flow.latent._latch_dim(X_train.shape[1])
    if "params" not in variables or "batch_stats" not in variables:
        raise KeyError("variables must hold both 'params' and 'batch_stats' (train.py:59-60)")

    engine = TrainEngine(flow, variables, int(X_train.shape[1]), _cdim(C_train), lr=optimizer.learning_rate, b1=optimizer.b1, b2=optimizer.b2,
                         eps=optimizer.eps, weight_decay=optimizer.weight_decay, nesterov=optimizer.nesterov, group=group)

    lib = _lib.load()
    acc = torch.zeros(2, dtype=torch.float64, device=dev)   # [-sum(lp), count]
    _, world = _dist_world(group)
    n_rows = X_train.shape[0]
    step_counts = _global_step_counts(n_rows, batch_size, dev, group, world)

    def metric_fn(vs, x, c) -> float:  # train.py:75-78, over all ranks' rows
        lp = flow.apply(vs, x, c)
        _lib.check(lib.zf_neg_sum(stream_ptr(), ptr(lp), lp.shape[0], ptr(acc)), "zf_neg_sum")
        acc[1] = lp.shape[0]
        _allreduce(acc, group, "sum")
        total, count = acc.tolist()
        return total / count

    def shuffled(t, epoch_seed, out):  # X_train[perm], train.py:104-108
        tt = t if t.ndim == 2 else t.reshape(-1, 1)
        _lib.check(lib.zf_permute_rows(stream_ptr(), ptr(tt), tt.shape[0], tt.shape[1], epoch_seed, ptr(out)),
                   "zf_permute_rows")
        return out

    history_train: List[float] = []
    history_test: List[float] = []
    epoch_iter = _progress_iter(epochs) if progress else range(epochs)

    X_perm = torch.empty_like(X_train)
    C_perm = None if C_train is None else torch.empty_like(C_train if C_train.ndim == 2 else C_train.reshape(-1, 1))
    best_epoch, best_variables = 0, engine.snapshot()
    for epoch in epoch_iter:
        # one pass over a fresh shuffle of the training set (train.py:104-117)
        epoch_seed = (int(seed) * 0x9E3779B97F4A7C15 + epoch + 1) & 0xFFFFFFFFFFFFFFFF  # fold_in(iter_key, epoch)
        shuffled(X_train, epoch_seed, X_perm)
        if C_train is not None:
            shuffled(C_train, epoch_seed, C_perm)
        xb = cb = None
        for k, lo in enumerate(range(0, n_rows, batch_size)):
            xb = X_perm[lo:lo + batch_size]
            cb = None if C_perm is None else C_perm[lo:lo + batch_size]
            engine.step(xb, cb, global_count=step_counts[k])

        # metrics on the last minibatch and on the test set (train.py:119-121)
        current = engine.variables()
        history_train.append(metric_fn(current, xb, cb))
        history_test.append(metric_fn(current, X_test, C_test))

        if not np.isfinite(history_train[-1]):  # train.py:123-126
            warnings.warn(f"epoch {epoch}: loss[train] not finite, abort training", RuntimeWarning)
            break
        if history_test[-1] <= history_test[best_epoch]:  # train.py:128-130
            best_epoch, best_variables = epoch, engine.snapshot()
        if _should_stop(history_test, epoch, warmup, patience):
            break

    return best_variables, best_epoch, history_train, history_test


def _global_step_counts(n_rows: int, batch_size: int, dev, group, world: int) -> List[int]:
    """Global number of rows of every minibatch step of an epoch.  Every rank must run the same number of steps
    (each step all-reduces statistics and gradients): raises ValueError on EVERY rank when the shards disagree."""
    local = [min(batch_size, n_rows - lo) for lo in range(0, n_rows, batch_size)]
    if world == 1:
        return local
    n = torch.tensor([len(local), -len(local)], dtype=torch.int64, device=dev)
    _allreduce(n, group, "max")
    n_max, n_min = int(n[0].item()), -int(n[1].item())
    if n_max != n_min:
        raise ValueError(f"data-parallel train(): ranks disagree on the number of minibatch steps per epoch "
                         f"({n_min} .. {n_max}); shard X_train so that ceil(rows / batch_size) is the same everywhere")
    t = torch.tensor(local, dtype=torch.int64, device=dev)
    _allreduce(t, group, "sum")
    return [int(v) for v in t.tolist()]


def _progress_iter(epochs: int):
    try:
        from tqdm.auto import tqdm as track
    except ModuleNotFoundError:  # pragma: no cover
        from rich.progress import track
    return track(range(epochs))


def _should_stop(history_test: List[float], epoch: int, warmup: int, patience: int) -> bool:
    """Patience-window early stop (train.py:132-136): every `patience` epochs after the warm-up, stop unless the
    best test loss of the last window beats the best of the window before it.  Like the reference, a patience that
    rounds down to 0 (epochs < 20 with the default fraction) divides by zero here."""
    if epoch < warmup or epoch < 2 * patience or epoch % patience != 0:
        return False
    recent = np.min(history_test[-patience:])
    before = np.min(history_test[-2 * patience:-patience])
    return not recent < before
