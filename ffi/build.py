"""Build ffi/libzenflow_b200_xla.so (the XLA FFI adapter) where JAX is installed.

    python ffi/build.py            # needs `import jax` (jax.ffi.include_dir()) and the built libzenflow_b200.so

In the image this repository is developed in JAX is absent: the script then says so and exits 2 (nothing is
stubbed).  ``--check`` type-checks the adapter against tests/stubs/ instead (what tests/test_ffi_sources.py runs).
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ffi", "zenflow_b200_xla.cc")
OUT = os.path.join(ROOT, "ffi", "libzenflow_b200_xla.so")
NATIVE = os.path.join(ROOT, "zenflow_b200", "_native")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def check_cmd():
    return ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "stubs"),
            "-I" + os.path.join(ROOT, "include"), SRC]


def main(argv):
    if "--check" in argv:
        return subprocess.call(check_cmd())
    try:
        import jax
        inc = jax.ffi.include_dir()
    except Exception as e:  # noqa: BLE001
        print(f"ffi/build.py: JAX is not importable here ({e}); the adapter can only be built where jax.ffi exists",
              file=sys.stderr)
        return 2
    if not os.path.exists(os.path.join(NATIVE, "libzenflow_b200.so")):
        print("ffi/build.py: build libzenflow_b200.so first (python -c 'import __graft_entry__ as g; g.build()')", file=sys.stderr)
        return 2
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + inc, "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"), SRC,
           "-L" + NATIVE, "-lzenflow_b200", "-L" + os.path.join(os.path.dirname(CUDA_INC), "lib64"), "-lcudart",
           "-Wl,-rpath," + NATIVE, "-o", OUT]
    print(" ".join(cmd))
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
