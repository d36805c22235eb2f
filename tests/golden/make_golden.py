"""Generate the committed golden fixtures.

The reference itself cannot be imported here (jax/flax are not installable), so two kinds of fixture exist:
  * reference_kats.json  - the known-answer vectors the reference's own tests hold for this path, transcribed with
                           the test's file:line (paths relative to the reference repository);
  * oracle_*.npz         - fixed-seed inputs / parameters / float64-oracle outputs of small flows.  They freeze the
                           oracle (tests/test_golden.py fails if oracle/zenflow_oracle.py drifts) and give the GPU
                           tests an expected value that does not depend on importing the oracle's code path.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import zenflow_oracle as zo  # noqa: E402

KATS = {
    "_index": {"source": "tests/test_utils.py:53-68",
               "x": [-2, -1, -0.5, -0.1, 0.0, 0.1, 0.5, 1.0, 1.5], "xk": [-1, 0, 1], "idx": [0, 0, 0, 0, 1, 1, 1, 2, 2]},
    "_knots": {"source": "tests/test_utils.py:71-74", "dx": [0.25, 0.25, 0.25], "xk": [0, 0.25, 0.5, 0.75]},
    "identity_spline": {"source": "tests/test_utils.py:7-13", "x_linspace": [-1, 2, 10], "bins": 4, "atol": 1e-5},
    "roll": {"source": "tests/test_bijectors.py:168-176", "x": [[1, 5], [3, 4], [6, 2]], "z": [[5, 1], [4, 3], [2, 6]]},
    "chain_of_rolls": {"source": "tests/test_bijectors.py:179-188", "x": [[1, 2, 3], [4, 5, 6]], "z": [[2, 3, 1], [5, 6, 4]]},
    "shift_bounds_margin_0.01": {"source": "tests/test_bijectors.py:35-58", "x": [[1, 5], [3, 4], [6, 2]],
                                 "xmin": [0.975, 1.985], "xmax": [6.025, 5.015]},
    "chain_shiftbounds_roll": {"source": "tests/test_bijectors.py:191-206", "x": [[2.5, 2, 3], [1, 3.5, 4.5], [4, 5, 6]],
                               "y": [[0.0, 0.5, 0.0], [0.5, 0.0, 0.5], [1.0, 1.0, 1.0]]},
    "beta_logpdf_norm": {"source": "SURVEY 8a-19 / distributions.py:100-104", "minus_betaln_12_12": 16.602059876},
}

CASES = {
    "oracle_two_moons_cond": dict(D=2, C=1, K=16, layers=(32, 32), n=None, roll=1, M=96, seed=11),
    "oracle_dim5_k7": dict(D=5, C=3, K=7, layers=(24, 16), n=None, roll=1, M=64, seed=12),
    "oracle_dim16_k32": dict(D=16, C=0, K=32, layers=(16,), n=3, roll=2, M=48, seed=13),
}


def flat_params(v):
    out = {}
    for col in ("params", "batch_stats"):
        for name, sub in v[col].items():
            for k1, leaf in sub.items():
                if isinstance(leaf, dict):
                    for k2, arr in leaf.items():
                        out[f"{col}/{name}/{k1}/{k2}"] = np.asarray(arr, np.float32)
                else:
                    out[f"{col}/{name}/{k1}"] = np.asarray(leaf, np.float32)
    return out


def main():
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
    for name, cfg in CASES.items():
        rng = np.random.default_rng(cfg["seed"])
        ops = zo.make_chain(cfg["D"], cfg["K"], cfg["layers"], n_couplings=cfg["n"], roll_shift=cfg["roll"])
        x = rng.normal(0.3, 1.1, (cfg["M"], cfg["D"])).astype(np.float32)
        c = rng.uniform(0, 1, (cfg["M"], cfg["C"])).astype(np.float32) if cfg["C"] else np.zeros((cfg["M"], 0), np.float32)
        cc = c if cfg["C"] else None
        v = zo.init_variables(ops, cfg["D"], cfg["C"], cfg["seed"], weight_scale=1.5, randomize_bn=True)
        _, _, stats = zo.chain_forward(ops, v, x, cc, train=True)
        v["batch_stats"]["bijectors_0"] = stats["bijectors_0"]
        v64 = {k: {a: {b: ({p: np.asarray(q, np.float64) for p, q in e.items()} if isinstance(e, dict) else np.asarray(e, np.float64))
                       for b, e in f.items()} for a, f in t.items()} for k, t in v.items()}
        x64, c64 = x.astype(np.float64), (None if cc is None else cc.astype(np.float64))
        y, ld, _ = zo.chain_forward(ops, v64, x64, c64)
        lp, _ = zo.flow_log_prob(ops, v64, x64, c64)
        u = rng.beta(12, 12, (cfg["M"], cfg["D"])).astype(np.float32)
        xinv = zo.chain_inverse(ops, v64, u.astype(np.float64), c64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, c=c, u=u, y=y, log_det=ld, log_prob=lp, x_inverse=xinv,
                            cfg=json.dumps({k: (list(val) if isinstance(val, tuple) else val) for k, val in cfg.items()}),
                            **flat_params(v))
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
