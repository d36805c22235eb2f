"""Base distributions of the flow — host-side mirror of zenflow/distributions.py.

``log_prob`` runs the latent log-pdf kernel (the same device code the fused
``Flow.__call__`` pass ends with); ``sample`` draws inside the CUDA library with counter-based
Philox streams (csrc/zf_rng.cuh).  The reference samples with ``jax.random`` streams that cannot
be reproduced without JAX, so the samplers are checked statistically, as the reference's own
tests do (tests/test_distributions.py:39-81).
"""
from __future__ import annotations

from abc import ABC
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._chain import ChainSpec
from ._device import like_input

__all__ = ["Distribution", "Normal", "TruncatedNormal", "Beta", "Uniform"]


def _seed_of(rngkey) -> int:
    """A 64-bit seed from what callers pass as a PRNG key (int, or the words of a jax PRNGKey)."""
    arr = np.asarray(0 if rngkey is None else rngkey).reshape(-1)
    seed = 0
    for v in arr:
        seed = (seed * 0x9E3779B97F4A7C15 + int(v) + 1) & 0xFFFFFFFFFFFFFFFF
    return seed


class Distribution(ABC):
    """Distribution base class with infrastructure for lazy initialization (distributions.py:11-46)."""

    _kind: str = ""
    __dim: Optional[int] = None

    def log_prob(self, x):
        """x (N, D) -> summed per-dimension log-pdf (N,).  Latches ``dim`` on first use."""
        if self.__dim is None:
            self.__dim = x.shape[-1]
        return self._log_prob_impl(x)

    def _latch_dim(self, dim: int) -> None:
        if self.__dim is None:
            self.__dim = int(dim)

    @property
    def dim(self):
        return self.__dim

    def _native(self):
        """(zf_latent_kind, peakness) of this distribution."""
        return _lib.LATENT_KINDS[self._kind], 1.0

    def _log_prob_impl(self, x):
        kind, peak = self._native()
        spec = ChainSpec(x.shape[-1], 0)  # an empty chain: only the latent log-pdf tail runs
        return like_input(spec.log_prob(x, None, kind, peak), x)

    def sample(self, nsamples: int, rngkey=None, *, as_numpy: bool = False):
        """(nsamples, dim) draws: counter-based Philox streams keyed by (seed, row, column) inside the CUDA
        library (csrc/zf_rng.cuh); needs ``dim`` (latched by an earlier log_prob).

        Return type (one convention for every sampler of the package): there is no input array to mirror, so the
        draws stay where they are produced - a CUDA ``torch.Tensor`` - unless ``as_numpy=True`` asks for a host
        copy (the reference returns a jax array that numpy can consume directly)."""
        if self.dim is None:
            raise ValueError("dim is not set yet: call log_prob (or evaluate the flow) once before sampling")
        kind, peak = self._native()
        spec = ChainSpec(self.dim, 0)  # an empty chain: the sampler is the inverse pass's tile load
        x = spec.sample(int(nsamples), None, kind, peak, _seed_of(rngkey))
        return x.cpu().numpy() if as_numpy else x

    def __repr__(self):
        """Return string representation."""
        return f"""{self.__class__.__name__}()"""


class Normal(Distribution):
    """Multivariate normal with mean 0.5 and standard deviation 0.1 (distributions.py:50-62)."""

    _kind = "normal"


class TruncatedNormal(Distribution):
    """Like :class:`Normal`, but truncated to the interval [0, 1] (distributions.py:65-78)."""

    _kind = "truncnorm"


class Beta(Distribution):
    """Multivariate symmetric beta distribution (distributions.py:81-116); density is exactly
    zero at the boundary.  ``peakness`` interpolates between uniform (1) and normal-like."""

    _kind = "beta"
    peakness: float

    def __init__(self, peakness: float = 12.0):
        if peakness < 1:
            raise ValueError("peakness must be at least 1")
        self.peakness = peakness

    def _native(self):
        return _lib.LATENT_KINDS["beta"], float(self.peakness)

    def __repr__(self):
        """Return string representation."""
        return f"{self.__class__.__name__}(peakness={self.peakness})"


class Uniform(Distribution):
    """Multivariate uniform distribution (distributions.py:119-126)."""

    _kind = "uniform"

