"""Parity evidence table: for every BASELINE config shape and every chain kernel, the achieved
max / p99.9 relative error of log_prob against the float64 oracle (with the float32 oracle's own error beside it),
forward / inverse errors, and the CHAIN-LEVEL bin-flip rate: the fraction of spline evaluations whose bin index
(zf_chain_bin_indices) differs from the float32 / float64 oracle chain's (bit-exact bins are only
guaranteed given identical raw parameters; the GEMM summation order upstream differs).

Run on a B200:  python scripts/parity_table.py > gpurun_out/r02_parity_table.md
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import zenflow_oracle as zo  # noqa: E402
from tests.helpers import product_chain, to64, trained_variables  # noqa: E402

CONFIGS = [
    # name (BASELINE config), D, C, K, layers, n_couplings, roll, M
    ("cfg1 two_moons", 2, 0, 16, (128, 128), None, 1, 10_000),
    ("cfg2 two_moons_conditional", 2, 1, 16, (128, 128), None, 1, 50_000),
    ("cfg3 deep_set flow", 2, 8, 16, (128,) * 6, None, 1, 1000),
    ("cfg4 bounded16", 16, 0, 32, (128, 128), 8, 2, 6000),
    ("cfg5 cond16", 16, 4, 32, (128, 128), 8, 2, 6000),
]


def oracle_bins(ops, v, x, c):
    out = []
    params, stats = v["params"], v["batch_stats"]
    for i, op in enumerate(ops):
        name = f"bijectors_{i}"
        if op["kind"] == "shift_bounds":
            x, _ = zo.shift_bounds_forward(x, dict(stats[name]), margin=op["margin"], bounds=op["bounds"], train=False)
        elif op["kind"] == "roll":
            x = zo.roll_forward(x, op["shift"])
        else:
            x, _, _, aux = zo.coupling_forward(x, c, params[name], stats[name], knots_=op["knots"], return_aux=True)
            out.append(np.asarray(aux["idx"]).reshape(x.shape[0], -1))
    return np.stack(out, axis=1)


def rel(err, truth):
    return err / (np.abs(truth) + 1e-30)


def main():
    import torch

    from zenflow_b200 import Flow, _lib
    from zenflow_b200._chain import ChainSpec
    from zenflow_b200.module import Scope

    print("# Parity table, round 2 (B200, %s)\n" % torch.cuda.get_device_name(0))
    print("Truth = float64 oracle (oracle/zenflow_oracle.py); `fp32 oracle` = the same restatement in float32, i.e. the "
          "reference's own arithmetic.  Relative errors are |got - truth| / |truth| of log_prob over the finite entries; "
          "`fwd y` / `inv x` are max absolute errors of Chain.__call__ / Chain.inverse outputs (unit-interval scale / "
          "data scale).  Bin flips: spline evaluations (event, coupling, dim) whose bin index from the CUDA chain "
          "differs from the oracle chain's, out of all evaluations.\n")
    print("| config | kernel | events | lp rel max | lp rel p99.9 | fp32 oracle rel max | fp32 oracle rel p99.9 | lp abs max | fwd y | inv x "
          "| bin flips vs fp32 oracle | vs fp64 oracle | fp32 oracle vs fp64 oracle |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, D, C, K, layers, nc, roll, M in CONFIGS:
        rng = np.random.default_rng(7)
        ops = zo.make_chain(D, K, layers, n_couplings=nc, roll_shift=roll)
        x = rng.normal(0.3, 1.2, (M, D)).astype(np.float32)
        c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
        v = trained_variables(ops, x, c, seed=1, weight_scale=1.0)
        v64 = to64(v)
        x64, c64 = x.astype(np.float64), None if c is None else c.astype(np.float64)
        lp64, _ = zo.flow_log_prob(ops, v64, x64, c64)
        lp32, _ = zo.flow_log_prob(ops, v, x, c)
        y64, _, _ = zo.chain_forward(ops, v64, x64, c64)
        u = rng.beta(12, 12, (M, D)).astype(np.float32)
        xi64 = zo.chain_inverse(ops, v64, u.astype(np.float64), c64)
        b32, b64 = oracle_bins(ops, v, x, c), oracle_bins(ops, v64, x64, c64)
        fin = np.isfinite(lp64) & (np.abs(lp64) < 1e30)
        r32 = rel(np.abs(lp32 - lp64), lp64)[fin]
        chain = product_chain(ops)
        flow = Flow(chain)
        fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
        kernels = [("tcgen05 (auto)", None), ("FFMA (simt)", "simt")]
        if D // 2 == 1:
            kernels.insert(1, ("tcgen05 single-tile (umma8)", "umma8"))
        for kname, impl in kernels:
            _lib.set_impl(impl)
            lp = np.asarray(flow.apply(fv, x, c), np.float64)
            y, _ = chain.apply(v, x, c)
            xi = chain.apply(v, u, c, method="inverse")
            spec = ChainSpec(D, C)
            chain._emit(spec, Scope(v))
            idx = spec.bin_indices(x, c).cpu().numpy().reshape(M, b32.shape[1], -1)
            r = rel(np.abs(lp - lp64), lp64)[fin]
            tot = idx.size
            print(f"| {name} | {kname} | {M} | {r.max():.2e} | {np.quantile(r, 0.999):.2e} | {r32.max():.2e} | "
                  f"{np.quantile(r32, 0.999):.2e} | {np.abs(lp - lp64)[fin].max():.2e} | {np.abs(y - y64).max():.2e} | "
                  f"{np.abs(xi - xi64).max():.2e} | {(idx != b32).sum()}/{tot} = {(idx != b32).mean():.2e} | "
                  f"{(idx != b64).sum()}/{tot} = {(idx != b64).mean():.2e} | {(b32 != b64).sum()}/{tot} = {(b32 != b64).mean():.2e} |")
        _lib.set_impl(None)
    print("\nGenerated by scripts/parity_table.py; gates live in tests/ (tests/helpers.py::assert_fp32_parity, "
          "tests/test_golden.py).")


if __name__ == "__main__":
    main()
