"""Developer microbenchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) for the shapes the chain kernel issues.
Needs the -DZF_TRACE variant:  python scripts/build_variant.py trace -DZF_TRACE   (container), then on the GPU box
python scripts/umma_rate.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "zenflow_b200", "_native", "libzenflow_b200_trace.so"))
lib.zf_debug_umma_rate.restype = C.c_int
lib.zf_debug_umma_rate.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_longlong)]
out = (C.c_longlong * 4)()
reps = 20
for ctas in (1, 148):
    for N in (48, 96, 128, 256):
        for mode, name in ((0, "3 products, one accumulator"), (1, "3 products, cross in 2nd accumulator"), (2, "1 product"),
                           (5, "3 products, 2 accumulators + concurrent tcgen05.ld")):
            if N == 256 and mode != 2 and mode != 0:
                continue
            rc = lib.zf_debug_umma_rate(N, reps, mode, ctas, out)
            assert rc == 0, rc
            n_mma = reps * 16 * (1 if mode & 2 else 3)
            print(f"ctas {ctas:3d} N {N:3d} {name:52s} issue {out[0]/n_mma:6.1f} cyc/mma   issue+drain {out[1]/n_mma:6.1f} cyc/mma")
