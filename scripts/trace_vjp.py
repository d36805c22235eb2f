"""Developer tool: per-phase clock64 timeline of one steady-state tile of the fused VJP kernel
(chain_umma_kernel<false, true>) inside a train step of the cond16 workload.
  build (container):  python scripts/trace_chain.py --build
  run (GPU box):      python scripts/trace_vjp.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zenflow_b200 import build as zb  # noqa: E402

os.environ["ZENFLOW_B200_NO_BUILD"] = "1"
zb.LIB_PATH = os.path.join(zb.LIB_DIR, "libzenflow_b200_trace.so")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from zenflow_b200 import _lib, Flow  # noqa: E402
from zenflow_b200 import bijectors as bi  # noqa: E402
from zenflow_b200._train import TrainEngine  # noqa: E402

M, D, Cc, K, n_c = 262144, 16, 4, 32, 8
mods = [bi.ShiftBounds()]
for i in range(n_c - 1):
    mods += [bi.NeuralSplineCoupling(knots=K, layers=(128, 128)), bi.Roll(2)]
mods.append(bi.NeuralSplineCoupling(knots=K, layers=(128, 128)))
flow = Flow(bi.Chain(mods))
variables = flow.init(0, np.zeros((1, D), np.float32), np.zeros((1, Cc), np.float32))
x = torch.rand(M, D, device="cuda")
c = torch.rand(M, Cc, device="cuda")
eng = TrainEngine(flow, variables, D, Cc, micro_batch=M)
lib = _lib.load()
buf = (C.c_longlong * 256)()
lib.zf_debug_trace_read.restype = C.c_int
eng.step(x, c)
torch.cuda.synchronize()
assert lib.zf_debug_trace_read(buf) == 0   # clears
eng.step(x, c)
torch.cuda.synchronize()
assert lib.zf_debug_trace_read(buf) == 0   # the last chain_umma launch of a step is the VJP kernel of the first coupling
t = np.array(buf[:]).reshape(4, 64)
t0 = t[0, 0]
for slot, name in ((0, "epilogue warp 0 (half 0)"), (1, "epilogue warp 4 (half 1)")):
    row = t[slot]
    row = row[row != 0]
    print(name)
    print("  t - t0 :", (row - t0).tolist())
    print("  deltas :", np.diff(row).tolist())
row = t[2]
row = row[row != 0]
print("MMA warp")
for i in range(0, len(row) - 4, 5):
    print(f"  unit start {row[i]-t0:7d}  issue end {row[i+1]-t0:7d}  waits: D-free {-row[i+2]:5d}  A-chunks {-row[i+3]:5d}  weights {-row[i+4]:5d}")
