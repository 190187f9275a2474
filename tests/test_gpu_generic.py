"""The generic (user-defined primitive + functor) path, SURVEY.md 8f N4: tests/cpp/generic_test.cu
(this repo's header templates, nvcc-compiled like a user's translation unit) against host brute
force, and against the REFERENCE's own generic path driven with the same user code
(oracle/_ref/ref_generic_driver, same tests/cpp/generic_prims.cuh): keys, sorted primitives,
deltas, leaves, nodes, root and per-ray results bit for bit."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "generic_test")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_generic_driver")


def test_generic_primitives_vs_host_brute_force():
    if not os.path.exists(BIN):
        pytest.fail("tests/cpp/generic_test not built (python -c 'import __graft_entry__ as g; g.build()')")
    r = subprocess.run([BIN, "200000", "65536"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASSED generic primitives" in r.stdout


@pytest.mark.parametrize("n_tris,n_rays", [(50000, 8192), (300000, 32768)])
def test_generic_primitives_vs_reference(tmp_path, n_tris, n_rays):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ref_generic_driver not built")
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir(); ref.mkdir()
    a = subprocess.run([BIN, str(n_tris), str(n_rays), str(ours)], capture_output=True, text=True, timeout=600)
    assert a.returncode == 0, a.stdout + a.stderr
    b = subprocess.run([REF, str(n_tris), str(n_rays), str(ref)], capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stdout + b.stderr
    for name in ("keys", "tris", "deltas", "leaves", "nodes", "root", "closest", "counts"):
        x = np.fromfile(ours / (name + ".bin"), np.uint32)
        y = np.fromfile(ref / (name + ".bin"), np.uint32)
        if name == "leaves":       # .zw are uninitialised in the reference (albvh.cuh:281-291)
            x, y = x.reshape(-1, 4)[:, :2], y.reshape(-1, 4)[:, :2]
        assert x.shape == y.shape, name
        assert np.array_equal(x, y), name
