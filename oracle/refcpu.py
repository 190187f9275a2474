"""ctypes binding of the reference-derived CPU checkers under oracle/_ref/ (built by
oracle/build_ref.sh from /root/reference; TEST INFRASTRUCTURE ONLY).

available() is False when the libraries were never built (e.g. no /root/reference)."""
import ctypes
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_P = ctypes.c_void_p
_lib = None
_hp = None
DRIVER = os.path.join(_DIR, "ref_driver")


def _load():
    global _lib, _hp
    if _lib is None and os.path.exists(os.path.join(_DIR, "libgrace_ref_cpu.so")):
        _lib = ctypes.CDLL(os.path.join(_DIR, "libgrace_ref_cpu.so"))
        _lib.ref_num_threads.restype = ctypes.c_int
        _lib.ref_sphere_hit.restype = ctypes.c_int
        _lib.ref_sphere_hit.argtypes = [_P, _P, _P, _P]
        _lib.ref_morton_key30.restype = ctypes.c_uint32
        _lib.ref_morton_key30.argtypes = [ctypes.c_uint32] * 3
        _lib.ref_morton_key63.restype = ctypes.c_uint64
        _lib.ref_morton_key63.argtypes = [ctypes.c_uint64] * 3
        _lib.ref_brute_hitcounts.argtypes = [_P, ctypes.c_long, _P, ctypes.c_long, _P]
        _lib.ref_brute_cumulative.argtypes = [_P, ctypes.c_long, _P, ctypes.c_long, _P]
    if _hp is None and os.path.exists(os.path.join(_DIR, "libchealpix.so")):
        _hp = ctypes.CDLL(os.path.join(_DIR, "libchealpix.so"))
        _hp.pix2vec_nest.argtypes = [ctypes.c_long, ctypes.c_long, _P]
    return _lib


def available():
    return _load() is not None


def driver_available():
    return os.path.exists(DRIVER)


def num_threads():
    return int(_load().ref_num_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sphere_hit(ray7, sphere4):
    r, s = _f32(ray7), _f32(sphere4)
    b2 = np.zeros(1, np.float32)
    d = np.zeros(1, np.float32)
    h = _load().ref_sphere_hit(r.ctypes.data_as(_P), s.ctypes.data_as(_P), b2.ctypes.data_as(_P),
                               d.ctypes.data_as(_P))
    return bool(h), float(b2[0]), float(d[0])


def morton_key30(x, y, z):
    return int(_load().ref_morton_key30(int(x), int(y), int(z)))


def morton_key63(x, y, z):
    return int(_load().ref_morton_key63(int(x), int(y), int(z)))


def brute_hitcounts(rays, spheres):
    r, s = _f32(rays).reshape(-1, 7), _f32(spheres)
    out = np.zeros(len(r), np.int32)
    _load().ref_brute_hitcounts(r.ctypes.data_as(_P), len(r), s.ctypes.data_as(_P), len(s), out.ctypes.data_as(_P))
    return out


def brute_cumulative(rays, spheres):
    r, s = _f32(rays).reshape(-1, 7), _f32(spheres)
    out = np.zeros(len(r), np.float32)
    _load().ref_brute_cumulative(r.ctypes.data_as(_P), len(r), s.ctypes.data_as(_P), len(s), out.ctypes.data_as(_P))
    return out


def pix2vec_nest(nside, pix):
    _load()
    if _hp is None:
        raise RuntimeError("libchealpix.so not built")
    pix = np.atleast_1d(pix)
    out = np.empty((len(pix), 3), np.float64)
    v = np.empty(3, np.float64)
    for i, p in enumerate(pix):
        _hp.pix2vec_nest(int(nside), int(p), v.ctypes.data_as(_P))
        out[i] = v
    return out
