"""Layout / accuracy probe of the 3xFP16 split product (zf_selftest_umma_f16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zenflow_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for N, K in [(128, 128), (96, 128), (48, 128), (16, 16), (64, 32)]:
    rng = np.random.default_rng(N + K)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    ref64 = A.astype(np.float64) @ B.astype(np.float64).T
    err32 = np.abs(A @ B.T - ref64).max()
    for variant in (0, 1, 2):
        out = torch.full((128, N), float("nan"), device="cuda")
        _lib.check(lib.zf_selftest_umma_f16(st, At.data_ptr(), Bt.data_ptr(), N, K, out.data_ptr(), variant))
        torch.cuda.synchronize()
        print(f"f16 N={N} K={K} variant {variant}: max err {np.abs(out.cpu().numpy() - ref64).max():.3e} (fp32 sgemm {err32:.2e})")
    out = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.zf_selftest_umma(st, At.data_ptr(), Bt.data_ptr(), N, K, out.data_ptr(), 0))
    print(f"tf32 N={N} K={K}: max err {np.abs(out.cpu().numpy() - ref64).max():.3e}")
