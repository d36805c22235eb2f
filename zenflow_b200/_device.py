"""Device-memory plumbing (PyTorch owns HBM buffers and streams; nothing else).

The product needs a CUDA device: every helper here raises if none is present.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


_cuda_seen = False


def require_cuda() -> torch.device:
    global _cuda_seen
    if not _cuda_seen:   # asked once: torch.cuda.is_available() goes through NVML on every call
        if not torch.cuda.is_available():
            raise RuntimeError("zenflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cuda_seen = True
    return torch.device("cuda", torch.cuda.current_device())


def to_device_f32(a, device: Optional[torch.device] = None) -> torch.Tensor:
    """numpy / torch / list -> contiguous float32 CUDA tensor (ints are cast like the
    reference does, bijectors.py:178-179)."""
    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter skips building a Stream object:
    this is on the path of every call, ~20 us otherwise)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def like_input(t: torch.Tensor, template):
    """Return results in the caller's currency: numpy in -> numpy out, torch in -> torch out."""
    if isinstance(template, torch.Tensor):
        return t
    return t.cpu().numpy()
