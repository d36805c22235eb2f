"""Shared helpers for the parity tests: build matching (oracle ops, product module, variables)."""
import numpy as np

from oracle import zenflow_oracle as zo


def product_chain(ops):
    from zenflow_b200 import bijectors as bi

    mods = []
    for op in ops:
        if op["kind"] == "shift_bounds":
            mods.append(bi.ShiftBounds(margin=op["margin"], bounds=op["bounds"]))
        elif op["kind"] == "roll":
            mods.append(bi.Roll(op["shift"]))
        else:
            mods.append(bi.NeuralSplineCoupling(knots=op["knots"], layers=op["layers"],
                                                act=getattr(bi, op.get("act", "swish"))))
    return bi.Chain(mods)


def to64(tree):
    if isinstance(tree, dict):
        return {k: to64(v) for k, v in tree.items()}
    return np.asarray(tree, np.float64)


def trained_variables(ops, x, c, seed=0, weight_scale=2.0):  # noqa: D401
    """Oracle-initialised variables with ShiftBounds/BatchNorm statistics set by one oracle
    train-mode pass over (x, c), BatchNorm params/biases randomised so every term matters."""
    D = x.shape[1]
    C = 0 if c is None else (1 if c.ndim == 1 else c.shape[1])
    v = zo.init_variables(ops, D, C, seed, weight_scale=weight_scale, randomize_bn=True)
    _, _, stats = zo.chain_forward(ops, v, x, c, train=True)
    # keep the randomised BN running stats (a train pass would move them only 1%); take ShiftBounds stats
    for name, st in stats.items():
        if any(k.startswith("xmin_") for k in st):
            v["batch_stats"][name] = st
    return v


def errs(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d = d[np.isfinite(d)]
    return float(d.max()) if d.size else 0.0


def assert_fp32_parity(got, truth64, oracle32, what, rtol=1e-5, atol=2e-5, slack=2.0):
    """Gate: as accurate as the reference's float32 arithmetic.  slack = allowed ratio of the worst
    sample's error to the float32 oracle's worst sample (2 for the FFMA kernels; 4 for the 3xTF32
    tensor-core kernel, whose operands carry 22 instead of 24 significant bits)."""
    got = np.asarray(got, np.float64)
    fin = np.isfinite(truth64) & (np.abs(truth64) < 1e30)
    np.testing.assert_array_equal(np.isfinite(got) & (np.abs(got) < 1e30), fin, err_msg=what)
    err = np.abs(got - truth64)[fin]
    ref = np.abs(np.asarray(oracle32, np.float64) - truth64)[fin]
    tol = rtol * np.abs(truth64[fin]) + atol
    q_got, q_ref = np.quantile(err / tol, 0.999), np.quantile(ref / tol, 0.999)
    # within tolerance wherever the reference's own float32 arithmetic is
    assert q_got <= max(1.0, 0.75 * slack * q_ref), f"{what}: 99.9% quantile of err/tol = {q_got:.2f} (fp32 oracle {q_ref:.2f})"
    assert err.max() <= slack * ref.max() + 1e-5, f"{what}: max err {err.max():.3e} vs fp32 oracle {ref.max():.3e}"


