"""CPU restatement of the Deep-Set conditioner Phi (TEST INFRASTRUCTURE, like the rest of oracle/: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it).

Follows /root/reference/examples/deep_set.ipynb (code cell 3, raw-JSON lines 138-160):
    NNBlock(out_dim, depth, width): depth x (Dense(width) -> swish), then Dense(out_dim)
    Phi: BatchNorm(use_running_average=not train) -> NNBlock(8, 3, 128) -> Dropout(0.3, deterministic=not train)
         -> sum_matrix @ x
and the BCOO sum matrix of ones built by ``preprocess`` (cell 1, lines 60-74), given here as its COO index list.
flax pieces restated from their documented behaviour exactly as in zenflow_oracle.py (Dense: x @ kernel + bias,
BatchNorm: momentum 0.99, eps 1e-5, biased batch variance; Dropout: keep with probability 1 - rate, scale by
1 / (1 - rate)).  The dropout keep-mask is an INPUT (jax's random stream is not reproducible without JAX).
Parity unpinned against a running JAX (not installable in this image), as stated in zenflow_oracle.py.
"""
import numpy as np

from oracle import zenflow_oracle as zo


def init_phi(in_dim=2, out_dim=8, depth=3, width=128, seed=0, randomize=True):
    rng = np.random.default_rng(seed)
    params = {"BatchNorm_0": {"scale": np.ones(in_dim, np.float32), "bias": np.zeros(in_dim, np.float32)}, "NNBlock_0": {}}
    stats = {"BatchNorm_0": {"mean": np.zeros(in_dim, np.float32), "var": np.ones(in_dim, np.float32)}}
    fan_in = in_dim
    for j, w in enumerate([width] * depth + [out_dim]):
        params["NNBlock_0"][f"Dense_{j}"] = {
            "kernel": (rng.standard_normal((fan_in, w)) / np.sqrt(fan_in)).astype(np.float32),   # flax default: lecun_normal
            "bias": (0.1 * rng.standard_normal(w)).astype(np.float32) if randomize else np.zeros(w, np.float32)}
        fan_in = w
    if randomize:
        params["BatchNorm_0"]["scale"] = rng.uniform(0.5, 1.5, in_dim).astype(np.float32)
        params["BatchNorm_0"]["bias"] = (0.1 * rng.standard_normal(in_dim)).astype(np.float32)
        stats["BatchNorm_0"]["mean"] = (0.2 * rng.standard_normal(in_dim)).astype(np.float32)
        stats["BatchNorm_0"]["var"] = rng.uniform(0.5, 1.5, in_dim).astype(np.float32)
    return {"params": params, "batch_stats": stats}


def phi_forward(variables, x, set_idx, row_idx, n_sets, *, train=False, dropout_mult=None):
    """deep_set.ipynb:152-160.  Returns (c (S, out), new batch_stats).  dropout_mult: (N, out) multipliers
    (0 or 1/(1-rate)) when train, else ignored."""
    p, st = variables["params"], variables["batch_stats"]
    bn_p, bn_s = p["BatchNorm_0"], st["BatchNorm_0"]
    h, new_mean, new_var = zo.batchnorm(x, bn_p["scale"], bn_p["bias"], bn_s["mean"], bn_s["var"], train)
    new_bn = {"mean": new_mean, "var": new_var}
    block = p["NNBlock_0"]
    n = len(block)
    for j in range(n - 1):
        h = zo.swish(zo.dense(h, block[f"Dense_{j}"]["kernel"], block[f"Dense_{j}"]["bias"]))
    h = zo.dense(h, block[f"Dense_{n - 1}"]["kernel"], block[f"Dense_{n - 1}"]["bias"])
    if train and dropout_mult is not None:
        h = h * dropout_mult.astype(h.dtype)
    c = np.zeros((n_sets, h.shape[1]), h.dtype)
    np.add.at(c, set_idx, h[row_idx])    # sum_matrix @ x with a matrix of ones
    return c, {"BatchNorm_0": new_bn if train else dict(bn_s)}


def coo_from_sizes(sizes, padded_rows=None):
    """The (set, row) index list ``preprocess`` builds (deep_set.ipynb:60-69): sets are consecutive row ranges."""
    sizes = np.asarray(sizes, int)
    set_idx = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    row_idx = np.arange(sizes.sum(), dtype=np.int32)
    return set_idx, row_idx
