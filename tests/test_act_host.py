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
