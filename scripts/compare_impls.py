"""Developer check: the chain kernel variants must agree bit for bit on the same inputs (a race or a missed barrier
shows up as a difference at scale).  python scripts/compare_impls.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zenflow_b200 import Flow
from zenflow_b200 import bijectors as bi

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_077
for (D, C, K, ncoup, shift) in [(2, 1, 16, 2, 1), (3, 0, 32, 3, 1), (2, 0, 16, 5, 1)]:
    mods = [bi.ShiftBounds()]
    for i in range(ncoup - 1):
        mods += [bi.NeuralSplineCoupling(knots=K, layers=(128, 128)), bi.Roll(shift)]
    mods.append(bi.NeuralSplineCoupling(knots=K, layers=(128, 128)))
    flow = Flow(bi.Chain(mods))
    g = torch.Generator(device="cuda").manual_seed(D * 100 + K)
    x = torch.rand(M, D, device="cuda", generator=g)
    c = torch.rand(M, C, device="cuda", generator=g) if C else None
    v = flow.init(0, x[:1].cpu().numpy(), None if c is None else c[:1].cpu().numpy())
    st = v["batch_stats"]["bijector"]["bijectors_0"]
    for i in range(D):
        st[f"xmin_{i}"] = np.array([-0.05], np.float32); st[f"xmax_{i}"] = np.array([1.05], np.float32)
    rng = np.random.default_rng(1)
    def jitter(t):
        return {k: jitter(a) if isinstance(a, dict) else (a + rng.normal(0, 0.3, a.shape).astype(np.float32) if a.ndim == 2 else a) for k, a in t.items()}
    v["params"] = jitter(v["params"])
    v = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(a).cuda(), v)
    out = {}
    for impl in ("default", "umma8", "simt"):
        if impl == "default": os.environ.pop("ZF_CHAIN_IMPL", None)
        else: os.environ["ZF_CHAIN_IMPL"] = impl
        lp = [flow.apply(v, x, c) for _ in range(3)]
        u = torch.rand(M, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) * 0.9 + 0.05
        xi = flow.bijector.apply({"params": v["params"]["bijector"], "batch_stats": v["batch_stats"]["bijector"]}, u, c, method="inverse")
        torch.cuda.synchronize()
        assert all(torch.equal(lp[0], t) for t in lp), f"{impl}: run-to-run difference"
        out[impl] = (lp[0], xi)
    same_lp = torch.equal(out["default"][0], out["umma8"][0])
    same_inv = torch.equal(out["default"][1], out["umma8"][1])
    d_simt = (out["default"][0] - out["simt"][0]).abs().max().item()
    print(f"D={D} C={C} K={K} couplings={ncoup} M={M}: two-tile vs single-tile log_prob bit-equal={same_lp}, inverse bit-equal={same_inv}; "
          f"max |tensor - FFMA| log_prob = {d_simt:.2e}; finite={torch.isfinite(out['default'][0]).all().item()}")
    assert same_lp and same_inv
print("ok")
