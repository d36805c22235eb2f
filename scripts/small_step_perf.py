"""Small-batch optimiser step (what the reference's train() runs: batch_size ~ 1000): eager launches against the
replayed CUDA graph of TrainEngine._step_graphed.  CUDA-event timings, one JSON line per case."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import zenflow_oracle as zo  # noqa: E402  (variables only)
from tests.helpers import product_chain  # noqa: E402
from zenflow_b200 import Flow, _lib  # noqa: E402
from zenflow_b200._train import TrainEngine  # noqa: E402

CASES = [("two_moons_conditional", 2, 1, 16, (128, 128), None, 1000),
         ("two_moons_conditional", 2, 1, 16, (128, 128), None, 16384),
         ("deep_set_flow", 2, 8, 16, (128,) * 6, None, 1000),
         ("cond16", 16, 4, 32, (128, 128), 8, 4096)]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1, help="index into CASES (default: all)")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--eager-only", action="store_true", help="for an ncu launch list of one step's kernels")
    args = ap.parse_args()
    for name, D, C, K, layers, ncoup, M in (CASES if args.case < 0 else [CASES[args.case]]):
        ops = zo.make_chain(D, K, layers, n_couplings=ncoup, roll_shift=1 if D == 2 else 2)
        v = zo.init_variables(ops, D, C, 1, weight_scale=1.0, randomize_bn=True)
        rng = np.random.default_rng(0)
        x = torch.from_numpy(rng.normal(0.2, 1.0, (M, D)).astype(np.float32)).cuda()
        c = torch.from_numpy(rng.uniform(0, 1, (M, C)).astype(np.float32)).cuda()
        out = {"case": name, "M": M}
        for graphs in ((False,) if args.eager_only else (False, True)):
            flow = Flow(product_chain(ops))
            flow.latent._latch_dim(D)
            eng = TrainEngine(flow, {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}, D, C)
            eng.use_graphs = graphs
            for _ in range(5):
                eng.step(x, c)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = args.steps
            e0.record()
            for _ in range(n):
                eng.step(x, c)
            e1.record()
            torch.cuda.synchronize()
            key = "graph" if graphs else "eager"
            out[key + "_ms"] = e0.elapsed_time(e1) / n
            out[key + "_launches"] = (_lib.launch_count() - n0) / n
        if "graph_ms" in out:
            out["speedup"] = out["eager_ms"] / out["graph_ms"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
