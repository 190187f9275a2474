"""ctypes loader for libgrace_b200.so (the C ABI declared in include/grace_b200.h).

There is no fallback: if the CUDA library is missing or cannot be loaded, importing
this module raises.  build() compiles it in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRACE_B200_LIB: development knob to load an A/B build of the same library (scripts/ab_variants.sh)
LIB_PATH = os.environ.get("GRACE_B200_LIB") or os.path.join(_HERE, "libgrace_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "grace_b200.h")


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see Makefile)."""
    cmd = ["make", "-C", _HERE, "-j8", "libgrace_b200.so"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def declared_symbols():
    """Every function name include/grace_b200.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grace_b200_[a-z0-9_]+)\s*\(", text)))


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(grace_b200 has no CPU fallback)" % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)
