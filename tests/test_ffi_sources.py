"""The jax.ffi adapter sources (ffi/) cannot run here (no JAX): check what can be checked.

* ffi/zenflow_b200_xla.cc compiles (-fsyntax-only -Wall -Werror) against a declaration-only stand-in of
  xla/ffi/api/ffi.h and the real include/zenflow_b200.h: every C-ABI call in it has the right arity and types;
* every C-ABI symbol it calls is exported by the built library;
* the pure-Python half of ffi/zenflow_jax.py (program encoding, leaf order) agrees with what the product's own host
  layer (zenflow_b200/bijectors.py -> ChainSpec) emits for the same chain.
"""
import importlib.util
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_adapter_type_checks_against_the_c_abi():
    build = _load("zf_ffi_build", "ffi/build.py")
    r = subprocess.run(build.check_cmd(), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_adapter_calls_only_exported_symbols():
    src = open(os.path.join(ROOT, "ffi", "zenflow_b200_xla.cc")).read()
    called = set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", src))
    header = open(os.path.join(ROOT, "include", "zenflow_b200.h")).read()
    declared = set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", header))
    assert called and called <= declared, called - declared
    so = os.path.join(ROOT, "zenflow_b200", "_native", "libzenflow_b200.so")
    if os.path.exists(so):
        out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
        exported = set(re.findall(r" T (zf_[a-z0-9_]+)", out))
        assert called <= exported, called - exported


def test_handler_symbols_match_python_targets():
    zj = _load("zf_ffi_jax", "ffi/zenflow_jax.py")
    src = open(os.path.join(ROOT, "ffi", "zenflow_b200_xla.cc")).read()
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+)", src))
    assert defined == set(zj._TARGETS.values())
    # every attribute a handler binds is passed by the python side under the same name
    for attr in set(re.findall(r'Attr<[^>]+>+\("(\w+)"\)', src)):
        assert re.search(rf"\b{attr}=", open(os.path.join(ROOT, "ffi", "zenflow_jax.py")).read()), attr


# duck-typed stand-ins of the reference's modules (bijectors.py:90,132,276,300): only the fields are read
def _ref_chain(dim, knots, layers, bounds, margin):
    def swish(x):
        return x
    mk = lambda name, **kw: type(name, (), kw)()
    bij = [mk("ShiftBounds", margin=margin, bounds=bounds)]
    for _ in range(dim):
        bij += [mk("NeuralSplineCoupling", knots=knots, layers=layers, act=swish), mk("Roll", shift=1)]
    return mk("Chain", bijectors=bij[:-1] if dim == 1 else bij)


def test_program_encoding_matches_the_host_layer():
    zj = _load("zf_ffi_jax", "ffi/zenflow_jax.py")
    D, K, layers = 3, 16, (128, 128)
    bounds = ((0, 0.0, 2.0), (1, 1.0, None), (2, None, 5.0))
    prog, bnds, margin = zj.encode_program(_ref_chain(D, K, layers, bounds, 0.1), D)
    assert prog[:1 + D] == [zj.OP_SHIFT_BOUNDS, zj.BOUND_BOTH, zj.BOUND_LOWER, zj.BOUND_UPPER]
    assert bnds == [0.0, 1.0, 0.0, 2.0, 0.0, 5.0] and margin == 0.1
    rest = prog[1 + D:]
    rec = [zj.OP_COUPLING, K, 2, 128, 128]
    assert rest == (rec + [zj.OP_ROLL, 1]) * D
    # the same chain through the product's host layer gives the same op kinds in the same order
    from zenflow_b200 import _lib
    assert (_lib.OP_SHIFT_BOUNDS, _lib.OP_ROLL, _lib.OP_COUPLING) == (zj.OP_SHIFT_BOUNDS, zj.OP_ROLL, zj.OP_COUPLING)
    assert (_lib.BOUND_NONE, _lib.BOUND_BOTH, _lib.BOUND_LOWER, _lib.BOUND_UPPER) == (0, 1, 2, 3)
    with pytest.raises(ValueError, match="out of bounds"):
        zj.encode_program(_ref_chain(2, K, layers, ((5, 0.0, 1.0),), 0.1), 2)
    with pytest.raises(ValueError, match="upper bound"):
        zj.encode_program(_ref_chain(2, K, layers, ((0, 1.0, 0.0),), 0.1), 2)


def test_leaf_order_matches_build_chain_from_leaves():
    zj = _load("zf_ffi_jax", "ffi/zenflow_jax.py")
    D, K, layers = 2, 16, (64,)
    chain = _ref_chain(D, K, layers, ((1, 0.0, 1.0),), 0.1)
    variables = {"params": {}, "batch_stats": {}}
    for i, b in enumerate(chain.bijectors):
        n = type(b).__name__
        if n == "ShiftBounds":
            variables["batch_stats"][f"bijectors_{i}"] = {"xmin_0": "xmin0", "xmax_0": "xmax0"}
        elif n == "NeuralSplineCoupling":
            variables["params"][f"bijectors_{i}"] = {
                "BatchNorm_0": {"scale": f"s{i}", "bias": f"b{i}"},
                "Dense_0": {"kernel": f"k0_{i}", "bias": f"b0_{i}"}, "Dense_1": {"kernel": f"k1_{i}", "bias": f"b1_{i}"}}
            variables["batch_stats"][f"bijectors_{i}"] = {"BatchNorm_0": {"mean": f"m{i}", "var": f"v{i}"}}
    leaves, is_stat, paths = zj.leaf_order(chain, variables, D)
    # ShiftBounds: packed xmin, xmax (fully bounded column 1 gets the ignored fill values), then per coupling
    # scale, bias, mean, var, kernel_0, bias_0, kernel_1, bias_1: the order of BuildChainFromLeaves (ffi/*.cc)
    assert leaves[0] == ["xmin0", 0.0] and leaves[1] == ["xmax0", 1.0]
    assert leaves[2:10] == ["s1", "b1", "m1", "v1", "k0_1", "b0_1", "k1_1", "b1_1"]
    assert leaves[10:18] == ["s3", "b3", "m3", "v3", "k0_3", "b0_3", "k1_3", "b1_3"]
    assert is_stat[:6] == [True, True, False, False, True, True] and not any(is_stat[6:10])
    assert paths[4] == ("batch_stats", 1, "BatchNorm_0", "mean")
    # count of trainable leaves per coupling = what BindGradients binds: 2 + 2 (n_hidden + 1)
    assert sum(1 for s in is_stat[2:10] if not s) == 2 + 2 * (len(layers) + 1)
