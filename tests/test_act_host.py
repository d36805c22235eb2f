"""NeuralSplineCoupling.act (bijectors.py:319) on the CPU side: the oracle's activations against torch's float64
implementations of the same jax.nn definitions, the host markers -> zf_act_kind, and the ffi program encoding."""
import importlib.util
import pathlib

import numpy as np
import pytest
import torch

from oracle import torch_oracle as to
from oracle import zenflow_oracle as zo

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("act", sorted(set(zo.ACTIVATIONS) - {"silu"}))
def test_oracle_activations_agree_with_torch(act):
    x = np.concatenate([np.linspace(-30, 30, 2001), [0.0, -0.0, 1e-8, -1e-8, 88.0, -88.0]])
    got = zo.ACTIVATIONS[act](x)
    ref = to._ACT[act](torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)  # atol: 1 + tanh(u) cancels for very negative x (gelu)
    got32 = zo.ACTIVATIONS[act](x.astype(np.float32))
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, ref, rtol=2e-6, atol=2e-6)


def test_markers_map_to_abi_kinds():
    from zenflow_b200 import _lib
    from zenflow_b200 import bijectors as bi

    header = (ROOT / "include" / "zenflow_b200.h").read_text()
    for name, kind in _lib.ACT_KINDS.items():
        assert f"ZF_ACT_{name.upper()} = {kind}" in header
        assert bi.NeuralSplineCoupling(act=getattr(bi, name))._act_kind == kind
        assert bi.NeuralSplineCoupling(act=name)._act_kind == kind
    assert bi.NeuralSplineCoupling()._act_kind == 0 and bi.NeuralSplineCoupling(act=bi.silu)._act_kind == 0

    def relu(x):  # a function named like the jax.nn one (what a reference user passes)
        return x
    assert bi.NeuralSplineCoupling(act=relu)._act_kind == _lib.ACT_KINDS["relu"]
    with pytest.raises(NotImplementedError, match="act="):
        bi.NeuralSplineCoupling(act=lambda x: x)
    with pytest.raises(RuntimeError, match="inside the CUDA kernels"):
        bi.tanh(1.0)


def test_ffi_program_carries_the_activation():
    spec = importlib.util.spec_from_file_location("zf_ffi_jax_act", ROOT / "ffi" / "zenflow_jax.py")
    zj = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(zj)
    from zenflow_b200 import _lib

    def gelu(x):
        return x
    mk = lambda name, **kw: type(name, (), kw)()
    chain = mk("Chain", bijectors=[mk("NeuralSplineCoupling", knots=32, layers=(128, 64), act=gelu)])
    prog, _, _ = zj.encode_program(chain, 4)
    assert prog == [zj.OP_COUPLING, 32 | _lib.ACT_KINDS["gelu"] << 16, 2, 128, 64]
    assert {k: v for k, v in zj.ACT_KINDS.items() if k != "silu"} == _lib.ACT_KINDS
    with pytest.raises(NotImplementedError):
        zj.encode_program(mk("Chain", bijectors=[mk("NeuralSplineCoupling", knots=8, layers=(8,), act=lambda x: x)]), 2)


@pytest.mark.parametrize("act", ["tanh", "sigmoid", "gelu", "elu", "softplus", "relu", "leaky_relu"])
def test_torch_oracle_with_activation_matches_numpy_and_central_differences(act):
    """The gradient truth of tests/test_gpu_act.py: torch float64 forward == numpy float64 oracle with the same `act`,
    and autograd == central differences of the numpy oracle (the reference's own method, tests/test_utils.py:38-47)."""
    rng = np.random.default_rng(3)
    D, C, M = 4, 2, 48
    ops = [dict(op, act=act) if op["kind"] == "coupling" else op for op in zo.make_chain(D, 8, (16, 16))]
    x = rng.normal(0.3, 1.0, (M, D))
    c = rng.uniform(0, 1, (M, C))
    v = zo.init_variables(ops, D, C, 1, weight_scale=2.0, randomize_bn=True)
    to64 = lambda t: {k: to64(u) for k, u in t.items()} if isinstance(t, dict) else np.array(t, np.float64)
    v64 = to64(v)
    lp_np, _ = zo.flow_log_prob(ops, v64, x, c, train=True)
    loss, grads, _, gc, lp_t = to.loss_and_grads(ops, v64, x, c)
    np.testing.assert_allclose(lp_t, lp_np, rtol=1e-10, atol=1e-10)
    # the activation matters: the default gives another answer on the same variables
    lp_swish, _ = zo.flow_log_prob(zo.make_chain(D, 8, (16, 16)), v64, x, c, train=True)
    assert np.abs(lp_swish - lp_np).max() > 1e-3

    def loss_np(vv, cc):
        lp, _ = zo.flow_log_prob(ops, vv, x, cc, train=True)
        return -lp.mean()

    h = 1e-6
    for name, layer, leaf, pos in [("bijectors_1", "Dense_2", "kernel", (3, 5)), ("bijectors_3", "Dense_0", "kernel", (1, 2)),
                                   ("bijectors_1", "Dense_1", "bias", (7,))]:
        vp, vm = to64(v64), to64(v64)
        vp["params"][name][layer][leaf][pos] += h
        vm["params"][name][layer][leaf][pos] -= h
        num = (loss_np(vp, c) - loss_np(vm, c)) / (2 * h)
        np.testing.assert_allclose(grads[name][layer][leaf][pos], num, rtol=5e-5, atol=1e-8)
    cp, cm = c.copy(), c.copy()
    cp[5, 1] += h
    cm[5, 1] -= h
    np.testing.assert_allclose(gc[5, 1], (loss_np(v64, cp) - loss_np(v64, cm)) / (2 * h), rtol=5e-5, atol=1e-9)
