"""Latency of flow.apply at notebook scale (BASELINE configs[0]: two_moons, 10k events, log_prob + sample):
numpy in / numpy out (what a reference user calls) and device-resident tensors; host profile of the numpy call."""
import cProfile
import io
import json
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from zenflow_b200 import Flow  # noqa: E402
from zenflow_b200.bijectors import rolling_spline_coupling  # noqa: E402


def wall(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    rng = np.random.default_rng(0)
    M = 10_000
    x = rng.normal(0.5, 0.5, (M, 2)).astype(np.float32)
    flow = Flow(rolling_spline_coupling(2))
    v = flow.init(0, x)
    vd = torch.utils._pytree.tree_map(lambda a: torch.as_tensor(a).cuda(), v)
    xd = torch.from_numpy(x).cuda()
    out = {"M": M}
    out["log_prob_numpy_ms"] = wall(lambda: flow.apply(v, x))
    out["log_prob_device_vars_numpy_x_ms"] = wall(lambda: flow.apply(vd, x))
    out["log_prob_device_ms"] = wall(lambda: flow.apply(vd, xd))
    out["sample_device_ms"] = wall(lambda: flow.apply(vd, M, method="sample"))
    out["sample_numpy_vars_ms"] = wall(lambda: flow.apply(v, M, method="sample"))
    print(json.dumps(out), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        flow.apply(vd, xd)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
