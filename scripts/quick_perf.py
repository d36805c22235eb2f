"""Developer timing of the kernels (not the official bench): CUDA events, warm-up, L2-sized inputs."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
if os.environ.get("ZF_LIB"):   # A/B against another build of the library (developer use)
    from zenflow_b200 import build as _zb
    _zb.LIB_PATH = os.path.abspath(os.environ["ZF_LIB"]); os.environ["ZENFLOW_B200_NO_BUILD"] = "1"
from zenflow_b200 import _lib, Flow
from zenflow_b200.utils import rqs_forward_raw, rqs_inverse_raw
from zenflow_b200 import bijectors as bi

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]

out = {}
for (M, d, K) in ([] if os.environ.get("QUICK_ONLY_FLOW") else [(8_000_000, 1, 16), (2_000_000, 8, 32)]):
    P = 3*K-1
    theta = torch.randn(M, d, P, device="cuda") * 1.0
    x = torch.rand(M, d, device="cuda")
    alg = 4*(d*P + 2*d + 1)*M
    best, med = timeit(lambda: rqs_forward_raw(x, theta, K))
    out[f"rqs_fwd_M{M}_d{d}_K{K}"] = dict(ms=best, med=med, GBs=alg/best/1e6, frac=alg/best/1e6/6541.5)
    alg = 4*(d*P + 2*d)*M
    best, med = timeit(lambda: rqs_inverse_raw(x, theta, K))
    out[f"rqs_inv_M{M}_d{d}_K{K}"] = dict(ms=best, med=med, GBs=alg/best/1e6, frac=alg/best/1e6/6541.5)
    del theta, x

def flow_case(D, C, K, layers, ncoup, shift, M):
    mods = [bi.ShiftBounds()]
    n = D if ncoup is None else ncoup
    for i in range(n - 1):
        mods += [bi.NeuralSplineCoupling(knots=K, layers=layers), bi.Roll(shift)]
    mods.append(bi.NeuralSplineCoupling(knots=K, layers=layers))
    flow = Flow(bi.Chain(mods))
    x = torch.rand(M, D, device="cuda"); c = torch.rand(M, C, device="cuda") if C else None
    v = flow.init(0, x[:1].cpu().numpy(), None if c is None else c[:1].cpu().numpy())
    st = v["batch_stats"]["bijector"]["bijectors_0"]
    for i in range(D):
        st[f"xmin_{i}"] = np.array([-0.05], np.float32); st[f"xmax_{i}"] = np.array([1.05], np.float32)
    v = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(a).cuda(), v)
    return flow, v, x, c

for name, cfg in [("cfg2", (2, 1, 16, (128, 128), None, 1, 1_000_000)), ("cfg4_1M", (16, 0, 32, (128, 128), 8, 2, 1_000_000))]:
    flow, v, x, c = flow_case(*cfg)
    best, med = timeit(lambda: flow.apply(v, x, c), n=3, warm=1)
    M = cfg[-1]
    out[f"logprob_{name}"] = dict(ms=best, med=med, events_per_s=M/best*1e3)
    u = torch.rand(M, cfg[0], device="cuda")*0.8+0.1
    best, med = timeit(lambda: flow.apply(v, u, c, method="inverse"), n=3, warm=1)
    out[f"inverse_{name}"] = dict(ms=best, med=med, events_per_s=M/best*1e3)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quick_perf.json", "w"), indent=1)
