"""Developer tool: per-phase clock64 timeline of one steady-state tile of chain_umma_kernel.

Builds a -DZF_TRACE variant of the library next to the product one (never loaded by the package),
runs log_prob on the default bench workload and prints cycle deltas per role.
  build (container):  python scripts/trace_chain.py --build
  run (GPU box):      python scripts/trace_chain.py
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zenflow_b200 import build as zb  # noqa: E402

TRACE_LIB = os.path.join(zb.LIB_DIR, "libzenflow_b200_trace.so")

if "--build" in sys.argv:  # same as: python scripts/build_variant.py trace -DZF_TRACE
    cmd = [zb.nvcc_path()] + [f for f in zb.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DZF_TRACE", "-o", TRACE_LIB] + zb._sources()
    subprocess.run(cmd, check=True)
    print("built", TRACE_LIB)
    sys.exit(0)

os.environ["ZENFLOW_B200_NO_BUILD"] = "1"
# the clock64 stamps live in the single-tile kernel (2-D flows default to the two-tiles-in-flight kernel)
os.environ.setdefault("ZF_CHAIN_IMPL", "umma8")
zb.LIB_PATH = os.path.abspath(os.environ["ZF_LIB"]) if os.environ.get("ZF_LIB") else TRACE_LIB
import numpy as np  # noqa: E402
import torch  # noqa: E402
from zenflow_b200 import _lib, Flow  # noqa: E402
from zenflow_b200 import bijectors as bi  # noqa: E402

D, Cc, K, layers, M = 2, 1, 16, (128, 128), 1_000_000
if "--d16" in sys.argv:
    D, Cc, K, layers, M = 16, 0, 32, (128, 128), 1_000_000
mods = [bi.ShiftBounds()]
n = D if D == 2 else 8
for i in range(n - 1):
    mods += [bi.NeuralSplineCoupling(knots=K, layers=layers), bi.Roll(1 if D == 2 else 2)]
mods.append(bi.NeuralSplineCoupling(knots=K, layers=layers))
flow = Flow(bi.Chain(mods))
x = torch.rand(M, D, device="cuda")
c = torch.rand(M, Cc, device="cuda") if Cc else None
v = flow.init(0, x[:1].cpu().numpy(), None if c is None else c[:1].cpu().numpy())
st = v["batch_stats"]["bijector"]["bijectors_0"]
for i in range(D):
    st[f"xmin_{i}"] = np.array([-0.05], np.float32)
    st[f"xmax_{i}"] = np.array([1.05], np.float32)
v = torch.utils._pytree.tree_map(lambda a: torch.from_numpy(a).cuda(), v)
for _ in range(3):
    flow.apply(v, x, c)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 256)()
lib.zf_debug_trace_read.restype = C.c_int
assert lib.zf_debug_trace_read(buf) == 0
t = np.array(buf[:]).reshape(4, 64)
t0 = t[0, 0]
names = {0: "epilogue warp 0 (half 0)", 1: "epilogue warp 4 (half 1)", 2: "MMA warp", 3: "another epilogue warp"}
for slot in (0, 1, 2):
    row = t[slot]
    row = row[row != 0]
    print(names[slot])
    if slot == 2:   # (unit start, issue end, -pending wait, -A wait, -weights wait) per unit
        for i in range(0, len(row) - 4, 5):
            print(f"  unit start {row[i]-t0:7d}  issue end {row[i+1]-t0:7d}  waits: D-free {-row[i+2]:5d}  A-chunks {-row[i+3]:5d}  weights {-row[i+4]:5d}")
        continue
    print("  t - t0 :", (row - t0).tolist())
    print("  deltas :", np.diff(row).tolist())
