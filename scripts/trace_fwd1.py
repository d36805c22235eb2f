"""Developer tool: clock64 timeline of one steady-state tile of chain_umma_kernel<false> running ONE coupling
(the shape of a train-step forward launch: 16-D conditional, K = 32).  Needs the -DZF_TRACE build (trace_chain.py --build)."""
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
from zenflow_b200 import _lib  # noqa: E402
from zenflow_b200 import bijectors as bi  # noqa: E402

M, D, Cc, K = 1 << 20, 16, 4, 32
chain = bi.Chain([bi.NeuralSplineCoupling(knots=K, layers=(128, 128))])
x = torch.rand(M, D, device="cuda")
c = torch.rand(M, Cc, device="cuda")
v = chain.init(0, x[:1].cpu().numpy(), c[:1].cpu().numpy())
v = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(np.asarray(a)).cuda(), v)
lib = _lib.load()
buf = (C.c_longlong * 256)()
lib.zf_debug_trace_read.restype = C.c_int
for _ in range(3):
    chain.apply(v, x, c)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); chain.apply(v, x, c); b.record(); torch.cuda.synchronize()
print("ms per launch pair (pack + kernel):", a.elapsed_time(b))
assert lib.zf_debug_trace_read(buf) == 0
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
