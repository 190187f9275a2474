"""The C-ABI library loads (no GPU needed) and exports every symbol the header declares."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grace-devel_b200"))


def test_library_exports_declared_symbols():
    import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.grace_b200_last_error.restype = ctypes.c_char_p
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = lib.grace_b200_create(ctypes.byref(h), 0)
    assert rc != 0 and b"no CPU fallback" in lib.grace_b200_last_error()


def test_oracle_is_not_imported_by_product():
    pkg = os.path.join(ROOT, "grace-devel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU oracle", ""), (dirpath, f)
