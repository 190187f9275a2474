"""Gadget-2 (type 1) loader: the product reader/writer against the numpy
restatement of the reference's reader (oracle/gadget.py, tests/helper/read_gadget.cuh:15-167)
and, on the GPU box, against the reference's own reader compiled unmodified
(oracle/_ref/ref_gadget_driver)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grace-devel_b200"))
from oracle import gadget as og  # noqa: E402
from util import clustered_spheres  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_gadget_driver")


def _lib():
    import _lib
    return _lib.load()


def product_write(path, s, n_other=0, mass_block=False):
    lib = _lib()
    fn = lib.grace_b200_write_gadget_f4
    fn.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
    s = np.ascontiguousarray(s, np.float32)
    assert fn(os.fsencode(path), s.ctypes.data, len(s), n_other, int(mass_block)) == 0


def product_info(path):
    lib = _lib()
    fn = lib.grace_b200_gadget_info
    fn.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    np6 = (ctypes.c_longlong * 6)()
    m6 = (ctypes.c_double * 6)()
    ng = ctypes.c_longlong(0)
    rc = fn(os.fsencode(path), np6, m6, ctypes.byref(ng))
    return rc, list(np6), list(m6), ng.value


@pytest.mark.parametrize("n_other,mass_block", [(0, False), (1234, False), (777, True)])
def test_writer_and_header_on_cpu(tmp_path, n_other, mass_block):
    """Host-only entry points (no GPU): the product writer's files are read back exactly by the
    restated reference reader, and the header parser agrees with files written by the oracle."""
    s = clustered_spheres(5000, seed=4)
    p1, p2 = str(tmp_path / "a.gdt"), str(tmp_path / "b.gdt")
    product_write(p1, s, n_other, mass_block)
    assert np.array_equal(og.read_gadget(p1).view(np.uint32), s.view(np.uint32))
    og.write_gadget(p2, s, n_other, mass_block)
    assert np.array_equal(og.read_gadget(p2).view(np.uint32), s.view(np.uint32))
    for p in (p1, p2):
        rc, np6, m6, ng = product_info(p)
        assert rc == 0 and ng == 5000 and np6[:2] == [5000, n_other]
        assert (m6[1] == 0.0) == mass_block
    assert os.path.getsize(p1) == os.path.getsize(p2)
    assert product_info(str(tmp_path / "missing"))[0] != 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_other,mass_block", [(4097, 0, False), (300000, 999, True), ((1 << 21) * 2 + 12345, 50000, False)])
def test_read_gadget_matches_reference_reader(gb, tmp_path, n, n_other, mass_block):
    import torch
    s = clustered_spheres(n, seed=6, n_halos=3)
    path = str(tmp_path / "snap.gdt")
    og.write_gadget(path, s, n_other, mass_block)
    want = og.read_gadget(path)
    got = gb.read_gadget(path)
    torch.cuda.synchronize()
    assert got.shape == (n, 4)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert gb.gadget_info(path)[2] == n
    # the loaded records feed the build directly (same stream, no synchronisation in between)
    tree = gb.Tree(n, 32)
    gb.build_tree(gb.read_gadget(path), tree)
    assert tree.n_leaves > 1
    if os.path.exists(REF):      # the reference's own reader, compiled unmodified
        out = str(tmp_path / "ref.bin")
        r = subprocess.run([REF, path, out], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        ref = np.fromfile(out, np.float32).reshape(-1, 4)
        assert json.loads(r.stdout.strip().splitlines()[-1])["n_gas"] == n
        assert np.array_equal(ref.view(np.uint32), got.cpu().numpy().view(np.uint32))


@pytest.mark.gpu
def test_read_gadget_errors(gb, tmp_path):
    import torch
    s = clustered_spheres(1000, seed=2)
    path = str(tmp_path / "snap.gdt")
    og.write_gadget(path, s)
    with pytest.raises(Exception):                       # buffer too small: GRACE_B200_ERANGE
        gb.read_gadget(path, torch.empty((10, 4), dtype=torch.float32, device="cuda"))
    with open(path, "r+b") as f:                         # truncated file
        f.truncate(os.path.getsize(path) - 100)
    with pytest.raises(Exception):
        gb.read_gadget(path)
    nogas = str(tmp_path / "nogas.gdt")
    og.write_gadget(nogas, s[:0], n_other=10)            # read_gadget.cuh:85-90 throws
    with pytest.raises(Exception):
        gb.read_gadget(nogas)
    with pytest.raises(Exception):
        gb.read_gadget(str(tmp_path / "missing.gdt"))
