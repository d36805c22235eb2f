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
    assert lib.zf_abi_version() == 1
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
