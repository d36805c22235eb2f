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
            mods.append(bi.NeuralSplineCoupling(knots=op["knots"], layers=op["layers"]))
    return bi.Chain(mods)


def to64(tree):
    if isinstance(tree, dict):
        return {k: to64(v) for k, v in tree.items()}
    return np.asarray(tree, np.float64)


def trained_variables(ops, x, c, seed=0, weight_scale=2.0):
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
