"""Train steps of the cond16 workload (BASELINE configs[4] shape) for timing sweeps and ncu launch lists.
usage: train_step_once.py [events] [warm steps] [micro_batch] [timed steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zenflow_b200 import Flow
from zenflow_b200 import bijectors as bi
from zenflow_b200._train import TrainEngine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mb = int(sys.argv[3]) if len(sys.argv) > 3 else (1 << 18)
timed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
D, C, K, n_c = 16, 4, 32, 8
mods = [bi.ShiftBounds()]
for i in range(n_c - 1):
    mods += [bi.NeuralSplineCoupling(knots=K, layers=(128, 128)), bi.Roll(2)]
mods.append(bi.NeuralSplineCoupling(knots=K, layers=(128, 128)))
flow = Flow(bi.Chain(mods))
variables = flow.init(0, np.zeros((1, D), np.float32), np.zeros((1, C), np.float32))
x = torch.rand(M, D, device="cuda"); c = torch.rand(M, C, device="cuda")
eng = TrainEngine(flow, variables, D, C, micro_batch=mb)
for _ in range(steps):
    eng.step(x, c)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(timed):
    lp = eng.step(x, c)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / timed
print(f"events {M} micro_batch {mb} step ms {ms:.3f} samples/s {M / ms * 1e3:.4g} loss {-lp.item() / M:.6f}")
