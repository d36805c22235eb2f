"""CPU checks of the C-ABI boundary: the library builds, loads and exports every symbol the
header declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "zenflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from zenflow_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert "zf_rqs_forward" in names and "zf_flow_log_prob" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zenflow_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in zenflow_b200/_lib.py"
    assert lib.zf_abi_version() == _lib.ABI_VERSION == 3
    assert os.path.dirname(_lib.library_path()).startswith(ROOT)  # in-tree, not site-packages


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors must equal the C structs (checked against a tiny C program)."""
    import subprocess
    import tempfile

    from zenflow_b200 import _lib

    src = r'''
#include <stdio.h>
#include "zenflow_b200.h"
int main(void){ printf("%zu %zu %zu %zu\n", sizeof(zf_shift_bounds), sizeof(zf_coupling), sizeof(zf_op), sizeof(zf_chain)); return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.ZfShiftBounds), ctypes.sizeof(_lib.ZfCoupling),
                     ctypes.sizeof(_lib.ZfOp), ctypes.sizeof(_lib.ZfChain)]


def test_argument_validation_without_gpu():
    """Bad arguments are rejected with a status + message before any CUDA call."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    rc = lib.zf_rqs_forward(None, None, None, 10, 1, 16, None, None, None)
    assert rc == 1 and b"null" in lib.zf_last_error().lower()
    rc = lib.zf_rqs_forward(None, None, None, 10, 0, 16, None, None, None)
    assert rc == 1
    ch = _lib.ZfChain()
    ch.dim, ch.cdim, ch.n_ops = 100, 0, 0
    assert lib.zf_chain_workspace_bytes(ctypes.byref(ch), 10) == 0
    assert b"dim" in lib.zf_last_error()
    with pytest.raises(_lib.ZenflowNativeError):
        _lib.check(1, "demo")


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under zenflow_b200/ may reference it."""
    pkg = os.path.join(ROOT, "zenflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "zenflow_oracle" not in text, f


def test_build_digest_is_path_independent(tmp_path):
    """The in-tree library ships to other machines with the sources: a copy of the tree under another path must see
    the same digest (it once hashed absolute paths, so every process on the GPU box rebuilt - and 8 ranks raced)."""
    import importlib.util
    import shutil

    from zenflow_b200 import build as zb

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = tmp_path / "relocated"
    shutil.copytree(os.path.join(root, "zenflow_b200", "csrc"), dst / "zenflow_b200" / "csrc")
    shutil.copytree(os.path.join(root, "include"), dst / "include")
    shutil.copy(os.path.join(root, "zenflow_b200", "build.py"), dst / "zenflow_b200" / "build.py")
    spec = importlib.util.spec_from_file_location("relocated_build", dst / "zenflow_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod._digest() == zb._digest()


def _fake_chain(D, C, K, hidden, n_couplings, bound_kind=0):
    """A zf_chain with well-formed descriptors and fake (non-null, aligned) device pointers: enough for the host-only
    planning entry zf_chain_workspace_bytes, which never dereferences them."""
    from zenflow_b200 import _lib

    keep = []
    ops = []
    sb = _lib.ZfShiftBounds()
    for i in range(D):
        sb.kind[i] = bound_kind
        sb.lo[i], sb.hi[i] = 0.0, 1.0
    sb.margin = 0.1
    sb.xmin = sb.xmax = 0x10000
    op = _lib.ZfOp(); op.kind = _lib.OP_SHIFT_BOUNDS; op.shift_bounds = ctypes.pointer(sb)
    ops.append(op); keep.append(sb)
    for j in range(n_couplings):
        cp = _lib.ZfCoupling()
        cp.knots = K
        cp.n_hidden = len(hidden)
        for i, w in enumerate(hidden):
            cp.hidden[i] = w
        cp.bn_scale = cp.bn_bias = cp.bn_mean = cp.bn_var = 0x10000
        for i in range(len(hidden) + 1):
            cp.kernel[i] = 0x20000
            cp.bias[i] = 0x30000
        op = _lib.ZfOp(); op.kind = _lib.OP_COUPLING; op.coupling = ctypes.pointer(cp)
        ops.append(op); keep.append(cp)
        if j + 1 < n_couplings:
            r = _lib.ZfOp(); r.kind = _lib.OP_ROLL; r.shift = 1
            ops.append(r)
    arr = (_lib.ZfOp * len(ops))(*ops)
    ch = _lib.ZfChain()
    ch.dim, ch.cdim, ch.n_ops = D, C, len(ops)
    ch.ops = ctypes.cast(arr, ctypes.POINTER(_lib.ZfOp))
    keep += [arr, ops]
    return ch, keep


def test_workspace_planning_is_host_only_and_validates():
    """zf_chain_workspace_bytes plans the packed layout on the host (no device needed): sizes are 16-byte multiples,
    grow with the chain, are independent of M, and malformed chains are refused with a message."""
    from zenflow_b200 import _lib

    lib = _lib.load()
    ch2, k2 = _fake_chain(2, 1, 16, (128, 128), 2)
    n2 = lib.zf_chain_workspace_bytes(ctypes.byref(ch2), 1000)
    assert n2 > 0 and n2 % 16 == 0
    assert lib.zf_chain_workspace_bytes(ctypes.byref(ch2), 10_000_000) == n2
    ch8, k8 = _fake_chain(16, 4, 32, (128, 128), 8)
    n8 = lib.zf_chain_workspace_bytes(ctypes.byref(ch8), 1000)
    assert n8 > 4 * n2
    # parameters + tf32 images + constant blocks of the tensor-core path: a few MB at most for the 16-D flow
    assert n8 < 64 << 20
    bad, kb = _fake_chain(2, 0, 16, (128,), 1)
    kb[1].knots = 0
    assert lib.zf_chain_workspace_bytes(ctypes.byref(bad), 10) == 0
    assert b"knots" in lib.zf_last_error()
    one, k1 = _fake_chain(1, 0, 16, (128,), 1)       # NeuralSplineCoupling needs dim >= 2 (bijectors.py:326)
    assert lib.zf_chain_workspace_bytes(ctypes.byref(one), 10) == 0
    assert b"dim" in lib.zf_last_error()


def test_concurrent_builds_compile_once(tmp_path):
    """One process per GPU imports the package at the same time; with a missing or stale library exactly one of them
    must compile (file lock), the others wait and find it fresh, and nobody sees a half-written file (atomic rename).
    A fake nvcc on PATH counts its invocations."""
    import shutil
    import stat
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = tmp_path / "tree"
    shutil.copytree(os.path.join(root, "zenflow_b200", "csrc"), dst / "zenflow_b200" / "csrc")
    shutil.copytree(os.path.join(root, "include"), dst / "include")
    shutil.copy(os.path.join(root, "zenflow_b200", "build.py"), dst / "zenflow_b200" / "build.py")
    bindir = tmp_path / "bin"
    bindir.mkdir()
    counter = tmp_path / "count.txt"
    fake = bindir / "nvcc"
    fake.write_text(
        "#!/bin/sh\n"
        f"echo run >> {counter}\n"
        "out=''\nwhile [ $# -gt 0 ]; do if [ \"$1\" = '-o' ]; then out=\"$2\"; fi; shift; done\n"
        "sleep 1\nprintf 'not-a-real-library' > \"$out\"\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    code = ("import importlib.util, sys\n"
            f"spec = importlib.util.spec_from_file_location('b', r'{dst / 'zenflow_b200' / 'build.py'}')\n"
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
            "p = m.build()\n"
            "assert open(p).read() == 'not-a-real-library'\n")
    env = dict(os.environ, PATH=f"{bindir}:{os.environ['PATH']}")
    procs = [subprocess.Popen([sys.executable, "-c", code], env=env) for _ in range(4)]
    assert all(p.wait(timeout=120) == 0 for p in procs)
    # exactly ONE build ran: one nvcc per translation unit plus the link step, not four times that
    n_units = len([f for f in os.listdir(os.path.join(root, "zenflow_b200", "csrc")) if f.endswith(".cu")])
    assert counter.read_text().count("run") == n_units + 1
