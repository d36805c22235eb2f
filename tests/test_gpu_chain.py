"""GPU parity of the fused whole-chain passes (zf_chain_forward / zf_chain_inverse /
zf_flow_log_prob) through the host API, against the oracle in fp32 and fp64."""
import numpy as np
import pytest

from oracle import zenflow_oracle as zo
from tests.helpers import assert_fp32_parity, errs, product_chain, to64, trained_variables

pytestmark = pytest.mark.gpu

# north_star: log_prob within rel 1e-5.  Truth is the float64 oracle.  Random-weight splines
# have a few ill-conditioned samples (bin slopes up to 1e5) on which fp32 arithmetic itself
# - the reference's arithmetic, i.e. the float32 oracle - misses rel 1e-5, so the gate is:
#   (a) 99.9% of the samples within rel 1e-5 + abs 2e-5 (floor for |lp| ~ 0 and for the summed
#       rounding of D latent terms and up to 8x8 log-det terms), and
#   (b) the worst sample no worse than 2x the float32 oracle's own worst sample (+1e-5).
LP_RTOL, LP_ATOL = 1e-5, 2e-5
Y_ATOL = 5e-6


def _data(M, D, C, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.3, 1.2, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
    return x, c


CONFIGS = [
    # name, D, C, K, layers, n_couplings, roll shift, M
    ("two_moons", 2, 0, 16, (128, 128), None, 1, 10_000),
    ("two_moons_conditional", 2, 1, 16, (128, 128), None, 1, 20_011),
    ("deep_set_flow", 2, 8, 16, (128,) * 6, None, 1, 1000),
    ("bounded16", 16, 0, 32, (128, 128), 8, 2, 3001),
    ("cond16", 16, 4, 32, (128, 128), 8, 2, 1500),
    ("one_hidden_k32", 3, 2, 32, (128,), 4, 1, 4099),      # tensor-core kernels with no hidden-layer GEMM
    ("three_hidden", 2, 0, 16, (128, 128, 128), 3, 1, 2500),
    # multi-dim couplings on the tensor-core kernel: > 16 conditioner inputs (first Dense on the FFMA pipe), K = 16 rows
    # in both theta buffers; and an odd number of transformed dims with a tensor-core first Dense of 4 inputs
    ("wide_cond24", 24, 8, 16, (128, 128), 3, 3, 1700),
    ("odd_dims6", 6, 1, 32, (128, 128), 4, 1, 2100),
    ("odd", 5, 3, 7, (64, 48), None, 1, 2000),
    ("wide", 3, 0, 4, (200,), None, 1, 333),
]


@pytest.fixture(params=["auto", "simt", "umma8"])
def chain_impl(request):
    """auto = tensor-core (tcgen05) kernel when the chain fits it, else the FFMA kernel; simt = force FFMA;
    umma8 = the single-tile tensor-core kernel (auto picks the two-tiles-in-flight kernel for single-dim couplings,
    the single-tile one otherwise); all fall back like auto when the chain does not fit."""
    from zenflow_b200 import _lib

    _lib.set_impl(None if request.param == "auto" else request.param)
    yield request.param
    _lib.set_impl(None)


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_chain_forward_logprob_inverse(cfg, chain_impl):
    from zenflow_b200 import Flow

    name, D, C, K, layers, ncoup, shift, M = cfg
    # the tensor-core kernel computes the conditioner with 3xTF32 split products: fp32-class, but
    # 22-bit operands and tensor-core accumulation leave ~3x the error of an fp32 FFMA chain
    tc = chain_impl != "simt"
    slack, atol_lp = (4.0, 5e-5) if tc else (3.0, LP_ATOL)
    ops = zo.make_chain(D, K, layers, n_couplings=ncoup, roll_shift=shift)
    x, c = _data(M, D, C, seed=len(name))
    v = trained_variables(ops, x, c, seed=1)
    chain = product_chain(ops)

    # --- Chain.__call__ eval
    y, ld = chain.apply(v, x, c, train=False)
    yo, ldo, _ = zo.chain_forward(ops, v, x, c)
    y64, ld64, _ = zo.chain_forward(ops, to64(v), x.astype(np.float64), None if c is None else c.astype(np.float64))
    e_or = errs(yo, y64)
    e_gpu = errs(y, y64)
    print(f"\n[{name}/{chain_impl}] y err gpu={e_gpu:.2e} oracle32={e_or:.2e}; ld err gpu={errs(ld, ld64):.2e} "
          f"oracle32={errs(ldo, ld64):.2e}")
    assert_fp32_parity(y, y64, yo, "y", rtol=0, atol=Y_ATOL, slack=2 * slack)
    assert_fp32_parity(ld, ld64, ldo, "log_det", atol=atol_lp, slack=2 * slack if len(layers) > 2 else slack)

    # --- Flow.__call__ (log_prob)
    flow = Flow(chain)
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    lp = flow.apply(fv, x, c)
    lp64, _ = zo.flow_log_prob(ops, to64(v), x.astype(np.float64), None if c is None else c.astype(np.float64))
    lpo, _ = zo.flow_log_prob(ops, v, x, c)
    print(f"[{name}] lp err gpu={errs(lp, lp64):.2e} oracle32={errs(lpo, lp64):.2e} |lp|max={np.abs(lp64).max():.1f}")
    assert lp.shape == (M,) and lp.dtype == np.float32
    assert_fp32_parity(lp, lp64, lpo, "log_prob", atol=atol_lp, slack=2 * slack if len(layers) > 2 else slack)

    # --- Chain.inverse on a given latent draw (parity mode of Flow.sample)
    u = np.random.default_rng(3).beta(12, 12, (M, D)).astype(np.float32)
    xi = chain.apply(v, u, c, method="inverse")
    xi64 = zo.chain_inverse(ops, to64(v), u.astype(np.float64), None if c is None else c.astype(np.float64))
    xio = zo.chain_inverse(ops, v, u, c)
    scale = np.abs(xi64).max()
    print(f"[{name}] inverse err gpu={errs(xi, xi64):.2e} oracle32={errs(xio, xi64):.2e} scale={scale:.1f}")
    assert_fp32_parity(xi, xi64, xio, "inverse", rtol=0, atol=(2 if tc else 1) * 5e-6 * max(1.0, scale), slack=slack)


def test_reference_kats_through_host_api():
    """tests/test_bijectors.py:168-206 (Roll / Chain golden vectors) on the CUDA path."""
    from zenflow_b200 import bijectors as bi

    x = np.array([[1, 5], [3, 4], [6, 2]])
    roll = bi.Roll()
    z, ld = roll.apply(roll.init(0, x, None), x, None, train=True)
    np.testing.assert_array_equal(z, [[5, 1], [4, 3], [2, 6]])
    np.testing.assert_array_equal(ld, np.zeros(3))
    np.testing.assert_array_equal(roll.apply({}, z, None, method="inverse"), x)

    x = np.array([[1, 2, 3], [4, 5, 6]])
    ch = bi.Chain([bi.Roll(), bi.Roll()])
    z, ld = ch.apply({}, x, None, train=False)
    np.testing.assert_array_equal(z, [[2, 3, 1], [5, 6, 4]])
    np.testing.assert_array_equal(ch.apply({}, z, None, method="inverse"), x)

    # ShiftBounds eval with the reference's golden statistics (test_ShiftBounds_1)
    sb = bi.ShiftBounds(margin=0.01)
    stats = {"batch_stats": {"xmin_0": np.array([0.975], np.float32), "xmax_0": np.array([6.025], np.float32),
                             "xmin_1": np.array([1.985], np.float32), "xmax_1": np.array([5.015], np.float32)}}
    x = np.array([[1, 5], [3, 4], [6, 2]])
    y, ld = sb.apply(stats, x, None)
    y_ref = np.column_stack([(x[:, 0] - 0.975) / (6.025 - 0.975), (x[:, 1] - 1.985) / (5.015 - 1.985)])
    np.testing.assert_allclose(y, y_ref, atol=5e-6)
    np.testing.assert_allclose(ld, np.log(1 / 5.05) + np.log(1 / 3.03), atol=5e-6)
    np.testing.assert_allclose(sb.apply(stats, y, None, method="inverse"), x, atol=6e-6)


def test_shift_bounds_bounded_variants_eval():
    """tests/test_bijectors.py:61-92 formulas, eval mode with oracle statistics."""
    from zenflow_b200 import bijectors as bi

    rng = np.random.default_rng(0)
    x = np.column_stack([2 * rng.uniform(size=50) - 1, rng.exponential(size=50) * 10 + 10,
                         1 - rng.exponential(size=50)]).astype(np.float32)
    bounds = [(0, -1, 1), (1, 10, None), (2, None, 1)]
    st = {}
    yo, ldo = zo.shift_bounds_forward(x, st, margin=0.0, bounds=bounds, train=True)
    sb = bi.ShiftBounds(margin=0.0, bounds=bounds)
    y, ld = sb.apply({"batch_stats": st}, x, None)
    np.testing.assert_allclose(y, yo, atol=2e-6)
    np.testing.assert_allclose(ld, ldo, rtol=1e-5, atol=1e-5)
    x2 = sb.apply({"batch_stats": st}, y, None, method="inverse")
    np.testing.assert_allclose(x2, zo.shift_bounds_inverse(yo, st, bounds=bounds), rtol=1e-5, atol=1e-5)


def test_latent_log_probs_on_device():
    """tests/test_distributions.py:11-74 through the latent kernel."""
    from zenflow_b200 import distributions as dist

    rng = np.random.default_rng(1)
    x = rng.uniform(size=(1000, 3)).astype(np.float32)
    x[0] = [0.0, 0.5, 1.0]
    x[1] = [-0.1, 0.5, 0.5]
    for d, kind in [(dist.Beta(), "beta"), (dist.Normal(), "normal"), (dist.TruncatedNormal(), "truncnorm"),
                    (dist.Uniform(), "uniform"), (dist.Beta(3.5), "beta")]:
        lp = d.log_prob(x)
        ref = zo.latent_log_prob(x.astype(np.float64), kind, getattr(d, "peakness", 12.0))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(lp[fin], ref[fin], rtol=2e-5, atol=2e-5)
        # flow.py:47 is applied by the kernel: -inf -> finfo.min
        assert (lp[~fin] == np.finfo(np.float32).min).all()
        assert d.dim == 3


def test_batch_permutation_is_bit_exact_at_full_size():
    """Samples are independent in eval mode: lp(x[perm]) == lp(x)[perm] bit-for-bit, at the
    two_moons_conditional bench size (1M) — catches any tile-boundary or scheduling bug."""
    import torch
    from zenflow_b200 import Flow

    D, C, M = 2, 1, 1_000_000
    ops = zo.make_chain(D)
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, (M, D)).astype(np.float32)
    c = rng.integers(0, 2, (M, 1)).astype(np.float32)
    v = trained_variables(ops, x[:5000], c[:5000])
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    xt, ct = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
    lp = flow.apply(fv, xt, ct)
    perm = torch.randperm(M, device="cuda")
    lp2 = flow.apply(fv, xt[perm], ct[perm])
    assert torch.equal(lp[perm], lp2)
    sub = np.r_[0:2048, M - 2048:M]
    lp64, _ = zo.flow_log_prob(ops, to64(v), x[sub].astype(np.float64), c[sub].astype(np.float64))
    lpo, _ = zo.flow_log_prob(ops, v, x[sub], c[sub])
    assert_fp32_parity(lp[sub].cpu().numpy(), lp64, lpo, "log_prob@1M", atol=5e-5, slack=4.0)
    # round trip inverse(forward(x)) ~ x (structural EPS mismatch allows ~1e-4)
    chain = flow.bijector
    y, _ = chain.apply(v, xt, ct)
    x2 = chain.apply(v, y, ct, method="inverse")
    inside = (y > 1e-3).all(1) & (y < 1 - 1e-3).all(1)
    assert float((x2 - xt)[inside].abs().max()) < 5e-4 * float(xt.abs().max())


def test_empty_and_tiny_batches():
    """Edge sizes: M = 0, 1 and a ragged tile through both chain kernels."""
    from zenflow_b200 import Flow

    ops = zo.make_chain(2)
    rng = np.random.default_rng(0)
    xs, cs = rng.normal(size=(300, 2)).astype(np.float32), rng.uniform(size=(300, 1)).astype(np.float32)
    v = trained_variables(ops, xs, cs)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    assert flow.apply(fv, xs[:0], cs[:0]).shape == (0,)
    for M in (1, 127, 129):
        lp = flow.apply(fv, xs[:M], cs[:M])
        lpo, _ = zo.flow_log_prob(ops, v, xs[:M], cs[:M])
        np.testing.assert_allclose(lp, lpo, rtol=2e-5, atol=2e-4)
    y, ld = flow.bijector.apply(v, xs[:0], cs[:0])
    assert y.shape == (0, 2) and ld.shape == (0,)


def test_nan_and_out_of_support_inputs():
    """flow.py:47: NaN log-probs become finfo.min; values outside the latent support give zero density."""
    from zenflow_b200 import Flow

    ops = zo.make_chain(2)
    rng = np.random.default_rng(1)
    xs = rng.normal(size=(512, 2)).astype(np.float32)
    v = trained_variables(ops, xs, None)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    bad = xs.copy()
    bad[0, 0] = np.nan
    bad[1, 1] = 1e6   # clipped to the cube's edge by ShiftBounds -> Beta density exactly 0
    bad[2, 0] = -1e6  # clipped to 0, then moved off the edge by the first spline (z >= EPS): finite, very small density
    lp = flow.apply(fv, bad)
    lpo, _ = zo.flow_log_prob(ops, v, bad)
    fmin = np.finfo(np.float32).min
    assert lp[0] == fmin and lp[1] == fmin
    assert (lpo[:2] == fmin).all()
    np.testing.assert_allclose(lp[2:], lpo[2:], rtol=2e-5, atol=2e-4)


def test_bounded16_at_full_size():
    """BASELINE configs[3]: 16-D flow, 8 couplings, K=32 at 16*2^20 events: size-independent properties
    (batch-permutation bit-exactness on slices, forward/inverse round trip) + oracle parity on the ends."""
    import torch
    from zenflow_b200 import Flow

    D, K, M = 16, 32, 16 * 2 ** 20
    ops = zo.make_chain(D, K, (128, 128), n_couplings=8, roll_shift=2)
    g = torch.Generator(device="cuda").manual_seed(0)
    xt = torch.rand(M, D, device="cuda", generator=g) * 2 - 0.5
    xs = xt[:4096].cpu().numpy()
    v = trained_variables(ops, xs, None, weight_scale=1.0)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    fv = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda(), fv)
    lp = flow.apply(fv, xt)
    assert lp.shape == (M,)
    sub = torch.cat([torch.arange(0, 3000), torch.arange(M - 3000, M), torch.randint(0, M, (3000,))]).cuda()
    assert torch.equal(flow.apply(fv, xt[sub]), lp[sub])          # tiling / scheduling independent
    ends = np.r_[0:1024, M - 1024:M]
    xe = xt[ends].cpu().numpy()
    lp64, _ = zo.flow_log_prob(ops, to64(v), xe.astype(np.float64))
    lp32, _ = zo.flow_log_prob(ops, v, xe)
    assert_fp32_parity(lp[ends].cpu().numpy(), lp64, lp32, "log_prob@16M", atol=5e-5, slack=4.0)
    chain = flow.bijector
    vb = {"params": fv["params"]["bijector"], "batch_stats": fv["batch_stats"]["bijector"]}
    y, _ = chain.apply(vb, xt[: 2 ** 20])
    x2 = chain.apply(vb, y, None, method="inverse")
    inside = (y > 1e-3).all(1) & (y < 1 - 1e-3).all(1)
    assert float((x2 - xt[: 2 ** 20])[inside].abs().max()) < 2e-3


def test_latent_samplers_moments():
    """tests/test_distributions.py:39-43,57-61,76-81: sample shapes, moments and support."""
    from zenflow_b200 import distributions as dist

    for d in (dist.Normal(), dist.TruncatedNormal(), dist.Beta(), dist.Uniform()):
        d._latch_dim(3)
        x = d.sample(20000, 0).cpu().numpy()
        assert x.shape == (20000, 3)
        np.testing.assert_allclose(x.mean(0), 0.5, atol=5e-2)
        if isinstance(d, (dist.Normal, dist.TruncatedNormal)):
            np.testing.assert_allclose(np.cov(x.T), 0.1 ** 2 * np.identity(3), atol=5e-2)
        if isinstance(d, dist.Beta):
            assert (x > 0).all() and (x < 1).all()
        if isinstance(d, dist.Uniform):
            assert x.min() >= 0 and x.max() < 1
        x2 = d.sample(16, 0).cpu().numpy()
        x3 = d.sample(16, 1).cpu().numpy()
        assert np.array_equal(x2, d.sample(16, 0).cpu().numpy()) and not np.array_equal(x2, x3)  # seeded
        assert np.array_equal(x2, x[:16])  # counter-based: a draw depends on (seed, row, column) only
    # distribution shape: Kolmogorov-Smirnov against the closed forms
    from scipy import stats as sps

    n = 200_000
    ref = {"normal": sps.norm(0.5, 0.1), "truncnorm": sps.truncnorm(-5, 5, 0.5, 0.1), "beta": sps.beta(12, 12),
           "uniform": sps.uniform(0, 1)}
    for d in (dist.Normal(), dist.TruncatedNormal(), dist.Beta(), dist.Uniform(), dist.Beta(2.5)):
        d._latch_dim(2)
        x = d.sample(n, 7).cpu().numpy()
        r = ref[d._kind] if getattr(d, "peakness", 12.0) == 12.0 else sps.beta(d.peakness, d.peakness)
        for j in range(2):
            ks = sps.kstest(x[:, j], r.cdf)
            assert ks.statistic < 4.0 / np.sqrt(n), (d, j, ks)
        assert abs(np.corrcoef(x.T)[0, 1]) < 0.01  # columns are independent streams


def test_flow_sample_and_steps():
    """tests/test_flow.py:7-29 + flow.py:80-95."""
    from zenflow_b200 import Flow
    from zenflow_b200.bijectors import ShiftBounds, rolling_spline_coupling
    from zenflow_b200.distributions import Uniform

    flow = Flow(ShiftBounds(), latent=Uniform())
    x = np.array([[3.0, 2.0], [1.0, 4.0], [5.0, 6.0]], dtype=np.float32)
    variables = flow.init(0, x)
    log_prob, variables = flow.apply(variables, x, train=True, mutable=["batch_stats"])
    assert log_prob.shape == (3,)
    x2 = flow.apply(variables, 1000, method="sample").cpu().numpy()
    assert x2.shape == (1000, 2)
    # one documented convention: int size -> device tensor unless as_numpy=True; numpy conditions -> numpy
    x2n = flow.apply(variables, 1000, method="sample", as_numpy=True)
    assert isinstance(x2n, np.ndarray) and np.array_equal(x2n, x2)
    assert isinstance(flow.latent.sample(5, 1, as_numpy=True), np.ndarray)
    # the fused pass equals "draw the latent, then bijector.inverse" (flow.py:76-77) bit for bit
    u = flow.latent.sample(1000, 0)
    assert np.array_equal(flow.apply(variables, u, method="inverse").cpu().numpy(), x2)
    assert not np.array_equal(flow.apply(variables, 1000, method="sample", seed=1).cpu().numpy(), x2)
    assert x2[:, 0].min() >= 1 - 0.21 and x2[:, 0].max() <= 5 + 0.21 and x2[:, 1].min() >= 2 - 0.21
    c = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    flow2 = Flow(ShiftBounds(), latent=Uniform())
    v2 = flow2.init(0, x)
    _, v2 = flow2.apply(v2, x, c, train=True, mutable=["batch_stats"])
    assert flow2.apply(v2, c, method="sample").shape == (3, 2)
    with pytest.raises(ValueError):
        flow2.apply(v2, x, method="_steps")
    f3 = Flow(rolling_spline_coupling(2), latent=Uniform())
    v3 = f3.init(0, x)
    _, upd = f3.apply(v3, x, train=True, mutable=["batch_stats"])
    v3 = {"params": v3["params"], "batch_stats": upd["batch_stats"]}
    steps = f3.apply(v3, x, method="_steps")
    assert len(steps) == 4 and all(s.shape == (3, 2) for s in steps)
    back = f3.apply(v3, steps[-1], method="_steps", inverse=True)
    np.testing.assert_allclose(back[-1], x, rtol=2e-4, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("D,C,K,ncoup,shift", [(16, 4, 32, 3, 2), (2, 4, 16, 2, 1)])
def test_unnormalised_conditioning_features(D, C, K, ncoup, shift):
    """Conditioning features of wildly different scales (1e6: x - mean alone would overflow fp16; 1e-6; 3e4 with an offset).  BatchNorm makes the reference
    invariant to them; the tensor-core first Dense feeds (x - mean) * mul, not x - mean, to the fp16 hi / lo' split, so
    it neither overflows fp16 nor drops to subnormals: log_prob stays at the float32 oracle's accuracy."""
    from zenflow_b200 import Flow

    M = 3000
    ops = zo.make_chain(D, K, (128, 128), n_couplings=ncoup, roll_shift=shift)
    x, c = _data(M, D, C, seed=21)
    c = (c.astype(np.float64) * np.array([1e6, 1e-6, 1.0, 3e4]) + np.array([0.0, 0.0, -2.0, 1e3])).astype(np.float32)
    v = trained_variables(ops, x, c, seed=2)
    d = D // 2
    for st in v["batch_stats"].values():   # running statistics of the conditioning features = the data's
        if "BatchNorm_0" in st:
            st["BatchNorm_0"]["mean"][D - d:] = c.mean(0)
            st["BatchNorm_0"]["var"][D - d:] = c.var(0)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    lp = flow.apply(fv, x, c)
    lp64, _ = zo.flow_log_prob(ops, to64(v), x.astype(np.float64), c.astype(np.float64))
    lpo, _ = zo.flow_log_prob(ops, v, x, c)
    print(f"[scaled c, D={D}] lp err gpu={errs(lp, lp64):.2e} oracle32={errs(lpo, lp64):.2e}")
    assert np.isfinite(lp).all()
    assert_fp32_parity(lp, lp64, lpo, "log_prob", atol=5e-5, slack=4.0)


@pytest.mark.parametrize("D,C,K,layers", [(2, 1, 16, (128, 128)), (5, 2, 8, (32, 16)), (16, 0, 32, (128, 128))])
def test_steps_match_oracle_steps(D, C, K, layers):
    """Flow._steps (flow.py:80-95): every per-bijector intermediate, forward and inverse, against the oracle's
    chain_forward / chain_inverse(return_steps=True) - not only shapes and the round trip."""
    from zenflow_b200 import Flow

    M = 777
    ncoup = 3 if D == 16 else None
    ops = zo.make_chain(D, K, layers, n_couplings=ncoup, roll_shift=2 if D == 16 else 1)
    x, c = _data(M, D, C, seed=21)
    v = trained_variables(ops, x, c, seed=4)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    v64 = to64(v)
    x64, c64 = x.astype(np.float64), None if c is None else c.astype(np.float64)
    steps = flow.apply(fv, x, c, method="_steps")
    _, _, _, ref = zo.chain_forward(ops, v64, x64, c64, return_steps=True)
    _, _, _, ref32 = zo.chain_forward(ops, v, x, c, return_steps=True)
    assert len(steps) == len(ref) == len(ops)
    for i, (got, want, want32) in enumerate(zip(steps, ref, ref32)):
        assert got.shape == (M, D)
        if ops[i]["kind"] == "roll":   # a permutation of the previous step: bit-exact (tests/test_bijectors.py:168-188)
            np.testing.assert_array_equal(got, np.roll(steps[i - 1], ops[i]["shift"], axis=-1))
        e, e32 = np.abs(got - want).max(), np.abs(want32 - want).max()
        assert e <= 4 * e32 + Y_ATOL, f"forward step {i} ({ops[i]['kind']}): err {e:.2e} vs fp32 oracle {e32:.2e}"
    z = steps[-1]
    back = flow.apply(fv, z, c, method="_steps", inverse=True)
    _, refb = zo.chain_inverse(ops, v64, z.astype(np.float64), c64, return_steps=True)
    _, refb32 = zo.chain_inverse(ops, v, z, c, return_steps=True)
    assert len(back) == len(refb) == len(ops)
    scale = np.abs(x).max()
    for i, (got, want, want32) in enumerate(zip(back, refb, refb32)):
        e, e32 = np.abs(got - want).max(), np.abs(want32 - want).max()
        assert e <= 4 * e32 + 2e-5 * scale, f"inverse step {i}: err {e:.2e} vs fp32 oracle {e32:.2e}"


@pytest.mark.gpu
@pytest.mark.parametrize("offset_rows", [0, 1, 3])
def test_unaligned_device_inputs_and_tail_tiles(offset_rows):
    """The tensor-core kernel fetches full tiles of 16-byte-aligned inputs with bulk copies one tile ahead and reads
    everything else (unaligned bases, the ragged last tile) itself: both paths must give the same numbers.  Row
    offsets 1 and 3 of a 3-column tensor start 12 and 36 bytes into the allocation."""
    import torch

    from zenflow_b200 import Flow

    D, C, K, M = 3, 1, 16, 5 * 128 + 77
    ops = zo.make_chain(D, K, (128, 128), n_couplings=None, roll_shift=1)
    x, c = _data(M + 8, D, C, seed=5)
    v = trained_variables(ops, x, c, seed=2)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
    xs, cs = xd[offset_rows:offset_rows + M], cd[offset_rows:offset_rows + M]
    assert xs.data_ptr() % 16 == (12 * offset_rows) % 16
    lp = flow.apply(fv, xs, cs).cpu().numpy()
    ref = flow.apply(fv, xs.clone(), cs.clone()).cpu().numpy()      # aligned copies of the same rows
    np.testing.assert_array_equal(lp, ref)
    lpo, _ = zo.flow_log_prob(ops, v, x[offset_rows:offset_rows + M], c[offset_rows:offset_rows + M])
    lp64, _ = zo.flow_log_prob(ops, to64(v), x[offset_rows:offset_rows + M].astype(np.float64),
                               c[offset_rows:offset_rows + M].astype(np.float64))
    assert_fp32_parity(lp, lp64, lpo, "log_prob", atol=5e-5, slack=4.0)


@pytest.mark.gpu
@pytest.mark.parametrize("D,C,K,ncoup", [(2, 1, 16, 2), (3, 0, 32, 3)])
def test_two_tile_kernel_is_bit_equal_to_single_tile(D, C, K, ncoup):
    """Flows with single-dim couplings run two tiles in flight (phase-specialised warps, shared tensor memory); the
    arithmetic per event is the same as in the single-tile kernel, so the outputs must agree bit for bit - a race
    or a missed barrier between the roles would show up as a difference at this size.  Also run to run."""
    import torch

    from zenflow_b200 import Flow, _lib

    M = 300_077
    ops = zo.make_chain(D, K, (128, 128), n_couplings=ncoup, roll_shift=1)
    x, c = _data(M, D, C, seed=9)
    v = trained_variables(ops, x[:4096], None if c is None else c[:4096], seed=3)
    flow = Flow(product_chain(ops))
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    fv = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda(), fv)
    xd = torch.from_numpy(x).cuda()
    cd = None if c is None else torch.from_numpy(c).cuda()
    u = torch.rand(M, D, device="cuda") * 0.9 + 0.05
    out = {}
    for impl in ("default", "umma8"):
        _lib.set_impl(None if impl == "default" else impl)
        lps = [flow.apply(fv, xd, cd) for _ in range(3)]
        inv = flow.bijector.apply({"params": fv["params"]["bijector"], "batch_stats": fv["batch_stats"]["bijector"]}, u, cd,
                                  method="inverse")
        assert all(torch.equal(lps[0], t) for t in lps[1:])
        out[impl] = (lps[0], inv)
    _lib.set_impl(None)
    assert torch.equal(out["default"][0], out["umma8"][0])
    assert torch.equal(out["default"][1], out["umma8"][1])
