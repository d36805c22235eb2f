import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import zenflow_oracle as zo
from tests.helpers import product_chain, to64
from zenflow_b200 import Flow
from zenflow_b200._train import TrainEngine

for (D, C, K, layers, ncoup, roll, M) in [(3, 0, 5, (40,), None, 1, 515), (4, 2, 8, (16, 16), None, 1, 300), (2, 1, 16, (128,128), None, 1, 700)]:
    rng = np.random.default_rng(M)
    ops = zo.make_chain(D, K, layers, n_couplings=ncoup, roll_shift=roll)
    x = rng.normal(0.3, 1.0, (M, D)).astype(np.float32)
    c = rng.uniform(0, 1, (M, C)).astype(np.float32) if C else None
    v = zo.init_variables(ops, D, C, 2, weight_scale=1.5, randomize_bn=True)
    z, ld, stats, steps = zo.chain_forward(ops, to64(v), x.astype(np.float64), None if c is None else c.astype(np.float64), train=True, return_steps=True)
    lp64, _ = zo.flow_log_prob(ops, to64(v), x.astype(np.float64), None if c is None else c.astype(np.float64), train=True)
    flow = Flow(product_chain(ops)); flow.latent._latch_dim(D)
    fv = {"params": {"bijector": v["params"]}, "batch_stats": {"bijector": v["batch_stats"]}}
    eng = TrainEngine(flow, fv, D, C, micro_batch=128)
    lp_sum = eng.step(x, c, update=False)
    print("case", D, C, K, "loss gpu", -lp_sum.item()/M, "oracle", -lp64.mean())
    b = eng._bufs[M]
    # oracle steps index: after each op; group outputs correspond to ops indices of last roll in group
    gi = 0; op_i = 0
    idxs = []
    for i, op in enumerate(ops):
        if op["kind"] != "roll":
            idxs.append(i)
        else:
            idxs[-1] = i
    for gi, oi in enumerate(idxs):
        st = b["states"][gi].cpu().numpy()
        print("  group", gi, "ops idx", oi, "max|state diff|", np.abs(st - steps[oi]).max())
    print("  ld diff", np.abs(b["ld"].cpu().numpy() - ld).max())
    for g in eng.groups:
        if g["kind"] == "cp":
            print("  bmean", g["bmean"].cpu().numpy()[:4], "bvar", g["bvar"].cpu().numpy()[:4])
