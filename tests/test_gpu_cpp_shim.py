"""Runs the C++ drop-in API tests (tests/cpp/shim_tests.cpp: the reference's own test programs
re-expressed against include/grace) on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "shim_tests")


def test_cpp_shim_programs(gb):
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")])
    out = subprocess.run([BIN, "200000", "200"], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.count("PASSED") == 7 and "FAILED" not in out.stdout


def test_cpp_multi_gpu_layer(gb):
    """tests/cpp/mgpu_test.cpp: the multi-GPU C ABI (libgrace_b200_mgpu.so: one process, ncclCommInitAll, tree
    replicated, rays dealt in 32-aligned tiles, NCCL broadcast + gather) on every visible device against the
    single-GPU ABI: same tree, same column densities, same hit counts, bit for bit, for both ways of building."""
    exe = os.path.join(ROOT, "tests", "cpp", "mgpu_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "mgpu_test"])
    out = subprocess.run([exe, "20"], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0 and "PASSED" in out.stdout and "FAILED" not in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
