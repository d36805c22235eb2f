"""One train step of the cond16 workload (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zenflow_b200 import Flow
from zenflow_b200 import bijectors as bi
from zenflow_b200._train import TrainEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
D, C, K, n_c = 16, 4, 32, 8
mods = [bi.ShiftBounds()]
for i in range(n_c - 1):
    mods += [bi.NeuralSplineCoupling(knots=K, layers=(128, 128)), bi.Roll(2)]
mods.append(bi.NeuralSplineCoupling(knots=K, layers=(128, 128)))
flow = Flow(bi.Chain(mods))
variables = flow.init(0, np.zeros((1, D), np.float32), np.zeros((1, C), np.float32))
x = torch.rand(M, D, device="cuda"); c = torch.rand(M, C, device="cuda")
eng = TrainEngine(flow, variables, D, C)
for _ in range(steps):
    eng.step(x, c)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.step(x, c); b.record(); torch.cuda.synchronize()
print("step ms", a.elapsed_time(b))
