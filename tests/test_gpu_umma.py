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
