"""Developer probe: tcgen05 kind::f16 (bf16 x 2 split) with K-major / MN-major shared-memory operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zenflow_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for N, K in ((128, 64), (96, 32), (32, 128)):
    A = torch.randn(128, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    ref = (A.double() @ B.double().T)
    for flags in range(8):
        if (flags & 4) and not (flags & 3):
            continue
        out = torch.full((128, N), float("nan"), device="cuda")
        _lib.check(lib.zf_selftest_umma_bf16(torch.cuda.current_stream().cuda_stream, A.data_ptr(), B.data_ptr(), N, K, out.data_ptr(), flags), "probe")
        torch.cuda.synchronize()
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"N={N} K={K} A_mn={flags & 1} B_mn={(flags >> 1) & 1} swap={(flags >> 2) & 1}: rel err {err:.3e}")
