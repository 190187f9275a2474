"""CPU oracle for the GRACE ray-tracing hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (grace-devel_b200/, include/) never does.

Thin numpy/ctypes binding over oracle/grace_oracle.c (built by oracle/Makefile).
Every function here follows the reference file:line cited next to its C body.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgrace_oracle.so")

RAY_FLOATS = 7  # grace::Ray, include/grace/ray.h:5-10


def build(force=False):
    """Compile the C restatement (gcc, OpenMP).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "libgrace_oracle.so"] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    build()
    try:
        return ctypes.CDLL(_LIB_PATH)
    except OSError:
        build(force=True)
        return ctypes.CDLL(_LIB_PATH)


_lib = _load()

_c = ctypes
_P = ctypes.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


for _name, _res, _args in [
    ("orc_space_by_two_10bit", _c.c_uint32, [_c.c_uint32]),
    ("orc_space_by_two_21bit", _c.c_uint64, [_c.c_uint64]),
    ("orc_morton_key30", _c.c_uint32, [_c.c_uint32] * 3),
    ("orc_morton_key63", _c.c_uint64, [_c.c_uint64] * 3),
    ("orc_bounds", None, [_P, _c.c_long, _P, _P]),
    ("orc_morton_keys30", None, [_P, _c.c_long, _P, _P, _P]),
    ("orc_morton_keys63", None, [_P, _c.c_long, _P, _P, _P]),
    ("orc_sort_perm_u32", None, [_P, _c.c_long, _P]),
    ("orc_sort_perm_u64", None, [_P, _c.c_long, _P]),
    ("orc_deltas_euclid", None, [_P, _c.c_long, _P]),
    ("orc_deltas_sarea", None, [_P, _c.c_long, _P]),
    ("orc_deltas_xor32", None, [_P, _c.c_long, _P]),
    ("orc_deltas_xor64", None, [_P, _c.c_long, _P]),
    ("orc_sphere_hit", _c.c_int, [_P, _P, _P, _P]),
    ("orc_aabbs_hit", _c.c_int, [_P, _P]),
    ("orc_kernel_table", _P, []),
    ("orc_kernel_integral", _c.c_float, [_c.c_float, _c.c_float]),
    ("orc_trace", _c.c_int, [_P, _c.c_long, _P, _P, _P, _c.c_long, _c.c_int, _c.c_int,
                             _P, _P, _P, _P, _P, _P, _P]),
    ("orc_brute", None, [_P, _c.c_long, _P, _c.c_long, _c.c_int, _P, _P]),
    ("orc_num_threads", _c.c_int, []),
    ("orc_sort_by_distance", None, [_P, _P, _c.c_long, _c.c_long, _P, _P]),
    ("orc_one_to_many_rays", None, [_P, _P, _c.c_int, _c.c_long, _P]),
    ("orc_ray_dir_key", _c.c_uint32, [_P]),
    ("orc_pix2vec_nest", None, [_c.c_long, _c.c_long, _P]),
]:
    _fn = getattr(_lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
for _sfx in ("f32", "u32", "u64"):
    _fn = getattr(_lib, "orc_build_leaves_" + _sfx)
    _fn.restype = _c.c_long
    _fn.argtypes = [_P, _c.c_long, _c.c_int, _P]
    _fn = getattr(_lib, "orc_leaf_deltas_" + _sfx)
    _fn.restype = None
    _fn.argtypes = [_P, _c.c_long, _P, _P]
    _fn = getattr(_lib, "orc_build_nodes_" + _sfx)
    _fn.restype = None
    _fn.argtypes = [_P, _c.c_long, _P, _P, _P, _P]

_SFX = {np.dtype(np.float32): "f32", np.dtype(np.uint32): "u32", np.dtype(np.uint64): "u64"}


def num_threads():
    return int(_lib.orc_num_threads())


def space_by_two_10bit(x):
    return int(_lib.orc_space_by_two_10bit(int(x)))


def space_by_two_21bit(x):
    return int(_lib.orc_space_by_two_21bit(int(x)))


def morton_key30(x, y, z):
    return int(_lib.orc_morton_key30(int(x), int(y), int(z)))


def morton_key63(x, y, z):
    return int(_lib.orc_morton_key63(int(x), int(y), int(z)))


def kernel_table():
    p = _lib.orc_kernel_table()
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(51,)).copy()


def kernel_integral(b2, h):
    return float(_lib.orc_kernel_integral(float(b2), float(h)))


def bounds(spheres):
    s = _f32(spheres)
    lo = np.empty(3, np.float32)
    hi = np.empty(3, np.float32)
    _lib.orc_bounds(_ptr(s), len(s), _ptr(lo), _ptr(hi))
    return lo, hi


def morton_keys(spheres, bot, top, bits=30):
    s = _f32(spheres)
    bot = _f32(bot)
    top = _f32(top)
    if bits == 30:
        keys = np.empty(len(s), np.uint32)
        _lib.orc_morton_keys30(_ptr(s), len(s), _ptr(bot), _ptr(top), _ptr(keys))
    else:
        keys = np.empty(len(s), np.uint64)
        _lib.orc_morton_keys63(_ptr(s), len(s), _ptr(bot), _ptr(top), _ptr(keys))
    return keys


def sort_perm(keys):
    keys = np.ascontiguousarray(keys)
    perm = np.empty(len(keys), np.int32)
    if keys.dtype == np.uint32:
        _lib.orc_sort_perm_u32(_ptr(keys), len(keys), _ptr(perm))
    elif keys.dtype == np.uint64:
        _lib.orc_sort_perm_u64(_ptr(keys), len(keys), _ptr(perm))
    else:
        raise TypeError(keys.dtype)
    return perm


def deltas_euclid(spheres):
    s = _f32(spheres)
    d = np.empty(len(s) + 1, np.float32)
    _lib.orc_deltas_euclid(_ptr(s), len(s), _ptr(d))
    return d


def deltas_sarea(spheres):
    s = _f32(spheres)
    d = np.empty(len(s) + 1, np.float32)
    _lib.orc_deltas_sarea(_ptr(s), len(s), _ptr(d))
    return d


def deltas_xor(keys):
    keys = np.ascontiguousarray(keys)
    d = np.empty(len(keys) + 1, keys.dtype)
    fn = _lib.orc_deltas_xor32 if keys.dtype == np.uint32 else _lib.orc_deltas_xor64
    fn(_ptr(keys), len(keys), _ptr(d))
    return d


class Tree:
    """Host mirror of grace::Tree (include/grace/cuda/nodes.h:14-58)."""

    def __init__(self, nodes, leaves, root, max_per_leaf):
        self.nodes = nodes      # (L-1, 16) int32, reference int4[4] layout
        self.leaves = leaves    # (L, 4) int32 {first, count, 0, 0}
        self.root = int(root)
        self.max_per_leaf = int(max_per_leaf)

    @property
    def n_leaves(self):
        return len(self.leaves)


def build_leaves(deltas, max_per_leaf):
    deltas = np.ascontiguousarray(deltas)
    n = len(deltas) - 1
    leaves = np.zeros((n, 4), np.int32)
    fn = getattr(_lib, "orc_build_leaves_" + _SFX[deltas.dtype])
    L = fn(_ptr(deltas), n, int(max_per_leaf), _ptr(leaves))
    return leaves[:L].copy()


def leaf_deltas(leaves, deltas):
    deltas = np.ascontiguousarray(deltas)
    leaves = _i32(leaves)
    out = np.empty(len(leaves) + 1, deltas.dtype)
    fn = getattr(_lib, "orc_leaf_deltas_" + _SFX[deltas.dtype])
    fn(_ptr(leaves), len(leaves), _ptr(deltas), _ptr(out))
    return out


def build_nodes(leaves, spheres, ldeltas):
    leaves = _i32(leaves)
    s = _f32(spheres)
    ldeltas = np.ascontiguousarray(ldeltas)
    L = len(leaves)
    nodes = np.zeros((max(L - 1, 0), 16), np.int32)
    root = np.zeros(1, np.int32)
    fn = getattr(_lib, "orc_build_nodes_" + _SFX[ldeltas.dtype])
    fn(_ptr(leaves), L, _ptr(s), _ptr(ldeltas), _ptr(nodes), _ptr(root))
    return nodes, int(root[0])


def build_tree(spheres, deltas, max_per_leaf):
    """build_ALBVH, include/grace/cuda/kernels/albvh.cuh:986-1021."""
    n = len(spheres)
    if n <= max_per_leaf:
        raise ValueError("max_per_leaf must be less than the total number of primitives.")
    leaves = build_leaves(deltas, max_per_leaf)
    ld = leaf_deltas(leaves, deltas)
    nodes, root = build_nodes(leaves, spheres, ld)
    return Tree(nodes, leaves, root, max_per_leaf)


def trace(rays, spheres, tree, mode, offsets=None, total=0, with_stats=False):
    """Packet traversal, include/grace/cuda/kernels/bintree_trace.cuh:119-193.

    mode 0 -> int32 hit counts; mode 1 -> float32 cumulative column density;
    mode 2 -> (indices, integrals, distances) written at per-ray offsets.
    """
    rays = _f32(rays).reshape(-1, RAY_FLOATS)
    s = _f32(spheres)
    n_rays = len(rays)
    counts = np.zeros(n_rays, np.int32) if mode == 0 else None
    cum = np.zeros(n_rays, np.float32) if mode == 1 else None
    idx = integ = dist = None
    off = None
    if mode == 2:
        off = _i32(offsets)
        idx = np.empty(total, np.int32)
        integ = np.empty(total, np.float32)
        dist = np.empty(total, np.float32)
    stats = np.zeros(((n_rays + 31) // 32, 3), np.int64) if with_stats else None
    depth = _lib.orc_trace(_ptr(rays), n_rays, _ptr(s), _ptr(tree.nodes), _ptr(tree.leaves),
                           tree.n_leaves, tree.root, mode, _ptr(counts), _ptr(cum),
                           _ptr(off), _ptr(idx), _ptr(integ), _ptr(dist), _ptr(stats))
    out = {0: counts, 1: cum, 2: (idx, integ, dist)}[mode]
    if with_stats:
        return out, stats, depth
    return out


def trace_hitcounts(rays, spheres, tree):
    return trace(rays, spheres, tree, 0)


def trace_cumulative(rays, spheres, tree):
    return trace(rays, spheres, tree, 1)


def trace_hits(rays, spheres, tree):
    """trace_sph, include/grace/cuda/trace_sph.cuh:112-168: counts -> exclusive scan -> fill."""
    counts = trace(rays, spheres, tree, 0)
    offsets = np.zeros(len(counts), np.int32)
    if len(counts) > 1:
        offsets[1:] = np.cumsum(counts[:-1], dtype=np.int64).astype(np.int32)
    total = int(counts.sum())
    idx, integ, dist = trace(rays, spheres, tree, 2, offsets=offsets, total=total)
    return offsets, idx, integ, dist


def sort_by_distance(dist, offsets, idx, data):
    """include/grace/cuda/sort.cuh:100-131 (stable per-ray sort; returns new arrays)."""
    dist = _f32(dist).copy()
    idx = _i32(idx).copy()
    data = _f32(data).copy()
    off = _i32(offsets)
    _lib.orc_sort_by_distance(_ptr(dist), _ptr(off), len(off), len(dist), _ptr(idx), _ptr(data))
    return dist, idx, data


def brute_hitcounts(rays, spheres):
    """tests/tree_traversal/tree_traversal.cu:65-79."""
    rays = _f32(rays).reshape(-1, RAY_FLOATS)
    s = _f32(spheres)
    out = np.zeros(len(rays), np.int32)
    _lib.orc_brute(_ptr(rays), len(rays), _ptr(s), len(s), 0, _ptr(out), None)
    return out


def brute_cumulative(rays, spheres):
    rays = _f32(rays).reshape(-1, RAY_FLOATS)
    s = _f32(spheres)
    out = np.zeros(len(rays), np.float32)
    _lib.orc_brute(_ptr(rays), len(rays), _ptr(s), len(s), 1, None, _ptr(out))
    return out


def one_to_many_rays(origin, points):
    pts = _f32(points)
    o = _f32(origin)
    rays = np.empty((len(pts), RAY_FLOATS), np.float32)
    _lib.orc_one_to_many_rays(_ptr(o), _ptr(pts), pts.shape[1], len(pts), _ptr(rays))
    return rays


def ray_dir_keys(rays):
    rays = _f32(rays).reshape(-1, RAY_FLOATS)
    return np.array([_lib.orc_ray_dir_key(_ptr(rays[i])) for i in range(len(rays))], np.uint32)


def pix2vec_nest(nside, pix):
    pix = np.atleast_1d(pix)
    out = np.empty((len(pix), 3), np.float64)
    v = np.empty(3, np.float64)
    for i, p in enumerate(pix):
        _lib.orc_pix2vec_nest(int(nside), int(p), _ptr(v))
        out[i] = v
    return out


def sort_spheres(spheres, bits=30, bot=None, top=None):
    """morton_keys{30,63}_sort_sph, include/grace/cuda/build_sph.cuh:41-82."""
    s = _f32(spheres)
    if bot is None:
        bot, top = bounds(s)
    keys = morton_keys(s, bot, top, bits)
    perm = sort_perm(keys)
    return s[perm].copy(), keys[perm].copy(), perm
