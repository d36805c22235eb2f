"""tcgen05 building blocks: the 3xTF32 split GEMM (A in tensor memory, B image in shared memory,
fp32 accumulator in tensor memory) must be fp32-accurate."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 128), (96, 128), (48, 128), (16, 8), (64, 32)])
def test_umma_3xtf32_gemm(N, K):
    from zenflow_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(N + K)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.zf_selftest_umma(torch.cuda.current_stream().cuda_stream, At.data_ptr(), Bt.data_ptr(), N, K,
                                    out.data_ptr()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref64 = A.astype(np.float64) @ B.astype(np.float64).T
    ref32 = A @ B.T
    err = np.abs(got - ref64).max()
    err32 = np.abs(ref32 - ref64).max()
    print(f"\nN={N} K={K}: 3xTF32 max err {err:.2e}, fp32 sgemm max err {err32:.2e}")
    assert err <= 4 * err32 + 2e-6


def test_umma_error_budget():
    """Diagnostic: where the 3xTF32 error comes from.  With operands that are exactly representable in
    tf32 (lo parts zero) the only error left is the tensor core's fp32 accumulation."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(7)
    N = K = 128

    def run(A, B):
        At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        out = torch.empty((128, N), device="cuda")
        _lib.check(lib.zf_selftest_umma(torch.cuda.current_stream().cuda_stream, At.data_ptr(), Bt.data_ptr(), N, K,
                                        out.data_ptr()))
        return out.cpu().numpy()

    def tf32(a):
        b = a.view(np.uint32).astype(np.uint64)
        b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
        return b.view(np.float32)

    A = np.abs(rng.normal(size=(128, K))).astype(np.float32)  # all positive: truncation shows up as bias
    B = np.abs(rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    for name, (a, b) in {"full fp32 operands": (A, B), "tf32-exact operands": (tf32(A.copy()), tf32(B.copy()))}.items():
        got = run(a, b)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        rel = (got - ref) / np.abs(ref)
        rel32 = ((a @ b.T) - ref) / np.abs(ref)
        print(f"\n{name}: 3xTF32 rel err max {np.abs(rel).max():.2e} mean {rel.mean():+.2e} rms {rel.std():.2e} | "
              f"fp32 sgemm max {np.abs(rel32).max():.2e} mean {rel32.mean():+.2e} rms {rel32.std():.2e}")
