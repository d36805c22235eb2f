"""Developer tool: build a variant of the library with extra -D flags next to the product one.
   python scripts/build_variant.py NAME -DZF_URING=5 ...   ->  zenflow_b200/_native/libzenflow_b200_NAME.so
   (use with ZF_LIB=... scripts/quick_perf.py; never loaded by the package)"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zenflow_b200 import build as zb
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(zb.LIB_DIR, f"libzenflow_b200_{name}.so")
cmd = [zb.nvcc_path()] + [f for f in zb.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags + ["-o", out] + zb._sources()
subprocess.run(cmd, check=True)
print("built", out)
