"""Base distributions of the flow — host-side mirror of zenflow/distributions.py.

``log_prob`` runs the latent log-pdf kernel (the same device code the fused
``Flow.__call__`` pass ends with); ``sample`` draws on the device.  The reference samples
with ``jax.random`` streams that cannot be reproduced without JAX, so samplers here are
checked statistically, as the reference's own tests do (tests/test_distributions.py:39-81).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._chain import ChainSpec
from ._device import like_input, require_cuda

__all__ = ["Distribution", "Normal", "TruncatedNormal", "Beta", "Uniform"]


def _generator(rngkey, device) -> torch.Generator:
    g = torch.Generator(device=device)
    if isinstance(rngkey, torch.Generator):
        return rngkey
    arr = np.asarray(0 if rngkey is None else rngkey).reshape(-1)
    seed = 0
    for v in arr:
        seed = (seed * 0x9E3779B1 + int(v)) & 0x7FFFFFFFFFFFFFFF
    g.manual_seed(seed)
    return g


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

    @abstractmethod
    def sample(self, nsamples: int, rngkey): ...

    def __repr__(self):
        """Return string representation."""
        return f"""{self.__class__.__name__}()"""


class Normal(Distribution):
    """Multivariate normal with mean 0.5 and standard deviation 0.1 (distributions.py:50-62)."""

    _kind = "normal"

    def sample(self, nsamples: int, rngkey=None):
        dev = require_cuda()
        z = torch.randn((nsamples, self.dim), generator=_generator(rngkey, dev), device=dev)
        return 0.5 + 0.1 * z


class TruncatedNormal(Distribution):
    """Like :class:`Normal`, but truncated to the interval [0, 1] (distributions.py:65-78)."""

    _kind = "truncnorm"

    def sample(self, nsamples: int, rngkey=None):
        dev = require_cuda()
        g = _generator(rngkey, dev)
        # inverse-cdf draw of a standard normal truncated to [-5, 5]
        lo = 0.5 * (1 + torch.erf(torch.tensor(-5.0 / 2 ** 0.5, device=dev)))
        hi = 0.5 * (1 + torch.erf(torch.tensor(5.0 / 2 ** 0.5, device=dev)))
        u = torch.rand((nsamples, self.dim), generator=g, device=dev, dtype=torch.float64)
        z = torch.erfinv(2 * (lo + u * (hi - lo)) - 1) * 2 ** 0.5
        return (0.5 + 0.1 * z.clamp(-5, 5)).to(torch.float32)


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

    def sample(self, nsamples: int, rngkey=None):
        dev = require_cuda()
        g = _generator(rngkey, dev)
        # Beta(p, p) = G1 / (G1 + G2) with G ~ Gamma(p) (the construction jax.random.beta uses)
        conc = torch.full((nsamples, self.dim), float(self.peakness), device=dev)
        torch.manual_seed(int(g.initial_seed()) & 0x7FFFFFFF)
        g1 = torch._standard_gamma(conc)
        g2 = torch._standard_gamma(conc)
        return g1 / (g1 + g2)

    def __repr__(self):
        """Return string representation."""
        return f"{self.__class__.__name__}(peakness={self.peakness})"


class Uniform(Distribution):
    """Multivariate uniform distribution (distributions.py:119-126)."""

    _kind = "uniform"

    def sample(self, nsamples: int, rngkey=None):
        dev = require_cuda()
        return torch.rand((nsamples, self.dim), generator=_generator(rngkey, dev), device=dev)
