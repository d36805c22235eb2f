"""zenflow_b200 — the spline-coupling hot path of zenflow as hand-written sm_100a CUDA
behind the reference's FLAX-module API (``Flow``, the bijectors, the latent distributions,
``train``).  CUDA only: there is no CPU fallback."""

from .flow import Flow
from .train import train  # shadows the submodule attribute, exactly as zenflow/__init__.py does

__all__ = ("Flow", "train")
