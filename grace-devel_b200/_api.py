"""Host-side mirror of GRACE's SPH convenience API (include/grace/cuda/{build_sph,
trace_sph,gen_rays,sort}.cuh, cuda/nodes.h) over the grace_b200 C ABI.

Same names, argument meaning and error behaviour as the reference's C++ templates;
`thrust::device_vector<T>` becomes a contiguous CUDA torch tensor (PyTorch is used for
device memory and streams only):

    spheres  float32 [N, 4]   {x, y, z, h}            (float4)
    rays     float32 [R, 7]   {dx,dy,dz,ox,oy,oz,len} (grace::Ray, ray.h:5-10)
    keys     int32 / int64 [N] holding the uint32 / uint64 bit patterns
    deltas   float32 / int32 / int64 [N + 1]

std::invalid_argument in the reference -> ValueError here; CUDA failures -> RuntimeError.
Every call goes through libgrace_b200.so; nothing here computes on the CPU.
"""
import ctypes
import os
import math

import torch

from . import _lib
from ._dist import sharded_trace, tiles_of_rank, take_local, scatter_back  # noqa: F401

__all__ = [
    "Tree", "RaySortType", "Octants", "N_table", "kernel_integral_table",
    "morton_keys_sph", "morton_keys30_sort_sph", "morton_keys63_sort_sph",
    "euclidean_deltas_sph", "surface_area_deltas_sph", "XOR_deltas_sph", "ALBVH_sph",
    "trace_hitcounts_sph", "trace_cumulative_sph", "trace_stats_sph", "trace_ray_cost_sph", "trace_packet_profile_sph", "trace_sph", "trace_with_sentinels_sph", "trace_sorted_tiles",
    "sort_by_distance", "sort_by_key", "exclusive_scan",
    "min_vec3", "max_vec3", "min_max_x", "min_vec4", "max_vec4",
    "uniform_random_rays", "uniform_random_rays_single_octant", "one_to_many_rays",
    "plane_parallel_random_rays", "orthographic_projection_rays", "pinhole_camera_rays",
    "healpix_rays", "synth_gadget_spheres", "exclusive_segmented_scan",
    "weighted_exclusive_segmented_scan", "offsets_to_segments", "read_gadget", "write_gadget", "gadget_info", "context", "lib", "build_tree", "set_trace_mode", "set_trace_budget", "set_trace_pool", "set_hit_list_passes", "trace_balance_stats", "device_error", "sharded_trace", "tiles_of_rank", "take_local", "scatter_back",
]

_c = ctypes
_P = ctypes.c_void_p
lib = _lib.load()

GRACE_B200_EINVAL = 1
GRACE_B200_ERANGE = 3
N_table = 51  # cuda/trace_sph.cuh:22


class RaySortType:      # include/grace/types.h:47-51
    NoSort = 0
    DirectionSort = 1
    EndPointSort = 2


class Octants:          # include/grace/types.h:36-45
    PPP, PPM, PMP, PMM, MPP, MPM, MMP, MMM = 7, 6, 5, 4, 3, 2, 1, 0


class _TreeStruct(ctypes.Structure):
    _fields_ = [("d_nodes", _P), ("d_leaves", _P), ("d_root", _P),
                ("n_leaves", _c.c_int), ("max_per_leaf", _c.c_int)]


def _sig(name, args, res=_c.c_int):
    fn = getattr(lib, name)
    fn.argtypes = args
    fn.restype = res
    return fn


_sz = _c.c_size_t
_create = _sig("grace_b200_create", [_c.POINTER(_P), _c.c_int])
_destroy = _sig("grace_b200_destroy", [_P])
_last_error = _sig("grace_b200_last_error", [], _c.c_char_p)
_reserve = _sig("grace_b200_reserve", [_P, _sz])
_bounds = _sig("grace_b200_bounds_f4", [_P, _P, _sz, _P, _P])
_minmax = _sig("grace_b200_minmax_f4", [_P, _P, _sz, _P, _P])
_keys30 = _sig("grace_b200_morton_keys30_f4", [_P, _P, _sz, _P, _P, _P])
_keys63 = _sig("grace_b200_morton_keys63_f4", [_P, _P, _sz, _P, _P, _P])
_sort32 = _sig("grace_b200_sort_pairs_u32", [_P, _P, _P, _c.c_int, _sz, _c.c_int, _P, _P])
_sort64 = _sig("grace_b200_sort_pairs_u64", [_P, _P, _P, _c.c_int, _sz, _c.c_int, _P, _P])
_msort = _sig("grace_b200_morton_sort_f4", [_P, _P, _sz, _c.c_int, _P, _P, _P, _P])
_d_euclid = _sig("grace_b200_deltas_euclid_f4", [_P, _P, _sz, _P, _P])
_d_sarea = _sig("grace_b200_deltas_sarea_f4", [_P, _P, _sz, _P, _P])
_d_xor32 = _sig("grace_b200_deltas_xor32", [_P, _P, _sz, _P, _P])
_d_xor64 = _sig("grace_b200_deltas_xor64", [_P, _P, _sz, _P, _P])
_build = _sig("grace_b200_albvh_build_f4",
              [_P, _P, _sz, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.POINTER(_c.c_int), _P])
_TS = _c.POINTER(_TreeStruct)
_t_counts = _sig("grace_b200_trace_hitcounts_f4", [_P, _P, _sz, _P, _sz, _TS, _P, _P])
_t_cum = _sig("grace_b200_trace_cumulative_f4", [_P, _P, _sz, _P, _sz, _TS, _P, _P])
_t_hcount = _sig("grace_b200_trace_hits_count_f4",
                 [_P, _P, _sz, _P, _sz, _TS, _c.c_int, _P, _c.POINTER(_c.c_longlong), _P])
_t_hfill = _sig("grace_b200_trace_hits_fill_f4", [_P, _P, _sz, _P, _sz, _TS, _P, _P, _P, _P, _P])
_t_stats = _sig("grace_b200_trace_stats_f4", [_P, _P, _sz, _P, _sz, _TS, _c.POINTER(_c.c_longlong * 4), _P])
_sort_dist = _sig("grace_b200_sort_by_distance", [_P, _P, _P, _sz, _sz, _P, _P, _P])
_scan = _sig("grace_b200_exclusive_scan_i32", [_P, _P, _P, _sz, _P, _P])
_segscan = _sig("grace_b200_exclusive_segmented_scan_f32", [_P, _P, _sz, _P, _sz, _P, _P])
_wsegscan = _sig("grace_b200_weighted_exclusive_segmented_scan_f32", [_P, _P, _P, _P, _P, _sz, _sz, _P, _P])
_off2seg = _sig("grace_b200_offsets_to_segments", [_P, _P, _sz, _P, _sz, _P])
_gadget_info = _sig("grace_b200_gadget_info", [_c.c_char_p, _P, _P, _P])
_gadget_read = _sig("grace_b200_read_gadget_f4", [_P, _c.c_char_p, _P, _sz, _P, _P])
_gadget_write = _sig("grace_b200_write_gadget_f4", [_c.c_char_p, _P, _sz, _sz, _c.c_int])
_table = _sig("grace_b200_kernel_integral_table", [_c.POINTER(_c.c_int)], _c.POINTER(_c.c_double))


def _opt(name, args):
    return _sig(name, args) if hasattr(lib, name) else None


_F3 = _c.c_float * 3
_rays_uniform = _opt("grace_b200_uniform_random_rays",
                     [_P, _P, _sz, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_int,
                      _c.c_ulonglong, _P])
_rays_o2m = _opt("grace_b200_one_to_many_rays",
                 [_P, _P, _sz, _c.c_float, _c.c_float, _c.c_float, _P, _c.c_int, _c.c_int, _P, _P, _P])
_rays_pp = _opt("grace_b200_plane_parallel_random_rays",
                [_P, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_float, _c.c_ulonglong, _P])
_rays_ortho = _opt("grace_b200_orthographic_projection_rays",
                   [_P, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_float, _c.c_float, _P])
_rays_pinhole = _opt("grace_b200_pinhole_camera_rays",
                     [_P, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_float, _c.c_float, _P])
_rays_healpix = _opt("grace_b200_healpix_rays",
                     [_P, _P, _sz, _c.c_long, _c.c_long, _c.c_float, _c.c_float, _c.c_float,
                      _c.c_float, _P])
_synth = _opt("grace_b200_synth_gadget_f4", [_P, _P, _sz, _c.c_uint, _P])

_contexts = {}


def context(device=None):
    """The per-device grace_b200 context (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("grace_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    ctx = _contexts.get(dev)
    if ctx is None:
        h = _P()
        rc = _create(ctypes.byref(h), dev)
        if rc:
            raise RuntimeError(_last_error().decode())
        ctx = _contexts[dev] = h
    return ctx


def _check(rc):
    if rc == 0:
        return
    msg = _last_error().decode()
    if rc == GRACE_B200_EINVAL:
        raise ValueError(msg)           # std::invalid_argument in the reference
    if rc == GRACE_B200_ERANGE:
        raise OverflowError(msg)
    raise RuntimeError(msg)


def _stream():
    return _P(torch.cuda.current_stream().cuda_stream)


def _dp(t):
    return _P(t.data_ptr()) if t is not None else _P(0)


def _need(t, dtype, what, last=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise TypeError("%s must be a contiguous CUDA tensor of dtype %s" % (what, dtype))
    if t.device.index != torch.cuda.current_device():
        # the context, its workspace and the stream are those of the current device
        raise ValueError("%s lives on %s but the current CUDA device is %d" % (what, t.device, torch.cuda.current_device()))
    if last is not None and (t.dim() != 2 or t.shape[1] != last):
        raise TypeError("%s must have shape [n, %d]" % (what, last))
    return t


def _h3(v):
    return _F3(float(v[0]), float(v[1]), float(v[2]))


def reserve(nbytes):
    _check(_reserve(context(), nbytes))


def set_trace_mode(mode):
    """'packet' (default), 'ray' (per-ray traversal) or 'packet_ref' (the reference's
    schedule and slab arithmetic bit for bit); see include/grace_b200.h."""
    _check(_sig("grace_b200_set_trace_mode", [_P, _c.c_int])(
        context(), {"ray": 0, "packet": 1, "packet_ref": 2}[mode]))


def set_trace_budget(steps, eager=False):
    """Steps before a still-running packet may be split once no unclaimed packet is left
    (eager=True: split at `steps` regardless; 0: never)."""
    _check(_sig("grace_b200_set_trace_budget", [_P, _c.c_int])(context(), int(steps) | ((1 << 30) if eager else 0)))


def set_trace_pool(nbytes):
    """Workspace for the per-hit terms recorded by column-density tasks (0 = automatic)."""
    _check(_sig("grace_b200_set_trace_pool", [_P, _sz])(context(), int(nbytes)))


def set_hit_list_passes(passes):
    """1 (default): trace_sph records the hits during the counting traversal; 2: count, then fill (the reference's scheme)."""
    _check(_sig("grace_b200_set_hit_list_passes", [_P, _c.c_int])(context(), int(passes)))


def trace_balance_stats():
    """Diagnostic: work-stealing counters of the last hit-count / column-density / hit-list count call
    (overflow: the pool recording the hit lists ran dry and the fill call traversed again)."""
    out = (_c.c_int * 8)()
    _check(_sig("grace_b200_trace_balance_stats", [_P, _P, _P])(context(), out, _stream()))
    return dict(finished=out[0], tasks=out[1], chunks=out[5], robbed=out[6], overflow=out[7])


def device_error():
    """Error flag of the last trace launch (0 ok, 1 stack overflow, 2 runaway traversal)."""
    f = _c.c_int(0)
    _check(_sig("grace_b200_device_error", [_P, _c.POINTER(_c.c_int), _P])(context(), ctypes.byref(f), _stream()))
    return f.value


def kernel_integral_table():
    n = _c.c_int()
    p = _table(ctypes.byref(n))
    return [p[i] for i in range(n.value)]


class Tree:
    """grace::Tree, include/grace/cuda/nodes.h:14-58.

    nodes  int32 [n_nodes, 16]: the reference's 4 x int4 per node
           ([0]={left,right,first_leaf,last_leaf}, [1]=L box {bx,tx,by,ty},
            [2]=R box, [3]={L bz,L tz,R bz,R tz}; floats bit-cast)
    leaves int32 [n_leaves, 4]: {first primitive, count, 0, 0}
    root_index_ptr: int32 [1] device tensor.
    Like the reference constructor, storage is sized for N_leaves leaves and shrunk by
    ALBVH_sph after clustering (albvh.cuh:842-845).
    """

    def __init__(self, N_leaves, max_per_leaf=1, device=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.nodes = torch.empty((max(int(N_leaves) - 1, 0), 16), dtype=torch.int32, device=dev)
        self.leaves = torch.empty((int(N_leaves), 4), dtype=torch.int32, device=dev)
        self.root_index_ptr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.max_per_leaf = int(max_per_leaf)

    @property
    def n_leaves(self):
        return self.leaves.shape[0]

    def _struct(self):
        return _TreeStruct(self.nodes.data_ptr(), self.leaves.data_ptr(),
                           self.root_index_ptr.data_ptr(), self.leaves.shape[0], self.max_per_leaf)


# ----------------------------------------------------------------------------- bounds
def _minmax8(spheres):
    _need(spheres, torch.float32, "spheres", 4)
    out = torch.empty(8, dtype=torch.float32, device=spheres.device)
    _check(_minmax(context(), _dp(spheres), spheres.shape[0], _dp(out), _stream()))
    return out


def min_vec3(spheres):
    """cuda/util/extrema.cuh:502-513 (returns a device tensor [3])."""
    return _minmax8(spheres)[0:3]


def max_vec3(spheres):
    return _minmax8(spheres)[4:7]


def min_vec4(spheres):
    return _minmax8(spheres)[0:4]


def max_vec4(spheres):
    return _minmax8(spheres)[4:8]


def min_max_x(spheres):
    """cuda/util/extrema.cuh:189-230 -> (min x, max x) as Python floats."""
    mm = _minmax8(spheres).cpu()
    return float(mm[0]), float(mm[4])


# ----------------------------------------------------------------------------- keys
def morton_keys_sph(d_spheres, d_keys, bot=None, top=None):
    """cuda/build_sph.cuh:19-34.  d_keys int32 -> 30-bit keys, int64 -> 63-bit keys.
    With bot/top None the bounds of the centres are computed (morton.cuh:139-173)."""
    _need(d_spheres, torch.float32, "d_spheres", 4)
    n = d_spheres.shape[0]
    if d_keys.dtype not in (torch.int32, torch.int64) or d_keys.numel() != n:
        raise TypeError("d_keys must be int32 (30-bit) or int64 (63-bit) of length N")
    if (bot is None) != (top is None):
        raise ValueError("bot and top must be given together")
    ctx = context()
    if bot is None:
        b6 = torch.empty(6, dtype=torch.float32, device=d_spheres.device)
        _check(_bounds(ctx, _dp(d_spheres), n, _dp(b6), _stream()))
    else:
        b6 = torch.tensor([bot[0], bot[1], bot[2], top[0], top[1], top[2]], dtype=torch.float32).to(
            d_spheres.device)
    fn = _keys30 if d_keys.dtype == torch.int32 else _keys63
    _check(fn(ctx, _dp(d_spheres), n, _dp(b6), _dp(d_keys), _stream()))
    return b6


def _morton_sort(d_spheres, bits, bot, top, return_keys):
    _need(d_spheres, torch.float32, "d_spheres", 4)
    n = d_spheres.shape[0]
    keys = None
    if return_keys:
        keys = torch.empty(n, dtype=torch.int32 if bits == 30 else torch.int64, device=d_spheres.device)
    hb = _h3(bot) if bot is not None else None
    ht = _h3(top) if top is not None else None
    _check(_msort(context(), _dp(d_spheres), n, bits, hb, ht, _dp(keys), _stream()))
    return keys


def morton_keys30_sort_sph(d_spheres, bot=None, top=None, return_keys=False):
    """cuda/build_sph.cuh:41-58: 30-bit keys, spheres sorted in place (stable)."""
    return _morton_sort(d_spheres, 30, bot, top, return_keys)


def morton_keys63_sort_sph(d_spheres, bot=None, top=None, return_keys=False):
    """cuda/build_sph.cuh:64-82."""
    return _morton_sort(d_spheres, 63, bot, top, return_keys)


def sort_by_key(d_keys, d_values=None, key_bits=None, return_perm=False):
    """thrust::sort_by_key equivalent: stable, ascending (keys as unsigned), in place."""
    if d_keys.dtype not in (torch.int32, torch.int64):
        raise TypeError("keys must be int32/int64 bit patterns")
    n = d_keys.numel()
    vb = 0
    if d_values is not None:
        vb = d_values.element_size() * (d_values.numel() // max(n, 1))
    perm = torch.empty(n, dtype=torch.int32, device=d_keys.device) if return_perm else None
    is32 = d_keys.dtype == torch.int32
    bits = key_bits or (32 if is32 else 64)
    fn = _sort32 if is32 else _sort64
    _check(fn(context(), _dp(d_keys), _dp(d_values), vb, n, bits, _dp(perm), _stream()))
    return perm


# ----------------------------------------------------------------------------- deltas
def euclidean_deltas_sph(d_spheres, d_deltas):
    """cuda/build_sph.cuh:87-93."""
    _need(d_spheres, torch.float32, "d_spheres", 4)
    _need(d_deltas, torch.float32, "d_deltas")
    if d_deltas.numel() != d_spheres.shape[0] + 1:
        raise ValueError("d_deltas must hold N + 1 elements")
    _check(_d_euclid(context(), _dp(d_spheres), d_spheres.shape[0], _dp(d_deltas), _stream()))


def surface_area_deltas_sph(d_spheres, d_deltas):
    """cuda/build_sph.cuh:97-103."""
    _need(d_spheres, torch.float32, "d_spheres", 4)
    _need(d_deltas, torch.float32, "d_deltas")
    if d_deltas.numel() != d_spheres.shape[0] + 1:
        raise ValueError("d_deltas must hold N + 1 elements")
    _check(_d_sarea(context(), _dp(d_spheres), d_spheres.shape[0], _dp(d_deltas), _stream()))


def XOR_deltas_sph(d_morton_keys, d_deltas):
    """cuda/build_sph.cuh:107-114."""
    if d_morton_keys.dtype not in (torch.int32, torch.int64) or d_morton_keys.dtype != d_deltas.dtype:
        raise TypeError("keys and deltas must both be int32 (30-bit keys) or int64 (63-bit keys)")
    if d_deltas.numel() != d_morton_keys.numel() + 1:
        raise ValueError("d_deltas must hold N + 1 elements")
    fn = _d_xor32 if d_morton_keys.dtype == torch.int32 else _d_xor64
    _check(fn(context(), _dp(d_morton_keys), d_morton_keys.numel(), _dp(d_deltas), _stream()))


_DELTA_TYPE = {torch.float32: 0, torch.int32: 1, torch.int64: 2}


def ALBVH_sph(d_spheres, d_deltas, d_tree):
    """cuda/build_sph.cuh:118-124 -> build_ALBVH (albvh.cuh:986-1021).
    Raises ValueError if N <= max_per_leaf (albvh.cuh:795-799)."""
    _need(d_spheres, torch.float32, "d_spheres", 4)
    n = d_spheres.shape[0]
    if d_deltas.dtype not in _DELTA_TYPE or d_deltas.numel() != n + 1:
        raise TypeError("d_deltas must be float32/int32/int64 of length N + 1")
    if d_tree.leaves.shape[0] < n or d_tree.nodes.shape[0] < n - 1:
        raise ValueError("Tree was constructed for fewer than N leaves")
    L = _c.c_int(0)
    _check(_build(context(), _dp(d_spheres), n, _dp(d_deltas), _DELTA_TYPE[d_deltas.dtype],
                  d_tree.max_per_leaf, _dp(d_tree.nodes), _dp(d_tree.leaves),
                  _dp(d_tree.root_index_ptr), ctypes.byref(L), _stream()))
    # remove_empty_leaves: resize nodes to 4*(L-1) int4 and leaves to L (albvh.cuh:842-845);
    # clone so the N-sized construction capacity goes back to the allocator.
    d_tree.nodes = d_tree.nodes[: L.value - 1].clone()
    d_tree.leaves = d_tree.leaves[: L.value].clone()
    return d_tree


def build_tree(spheres, tree, low=None, high=None, key_bits=30):
    """tests/helper/tree.cuh:15-43: 30-bit keys + sort, Euclidean deltas, ALBVH
    (key_bits=63: the same recipe with morton_keys63_sort_sph)."""
    deltas = torch.empty(spheres.shape[0] + 1, dtype=torch.float32, device=spheres.device)
    (morton_keys30_sort_sph if key_bits == 30 else morton_keys63_sort_sph)(spheres, low, high)
    euclidean_deltas_sph(spheres, deltas)
    ALBVH_sph(spheres, deltas, tree)
    return tree


# ----------------------------------------------------------------------------- trace
def _trace_args(d_rays, d_spheres, d_tree):
    _need(d_rays, torch.float32, "d_rays", 7)
    _need(d_spheres, torch.float32, "d_spheres", 4)
    return (context(), _dp(d_rays), d_rays.shape[0], _dp(d_spheres), d_spheres.shape[0],
            ctypes.byref(d_tree._struct()))


def trace_hitcounts_sph(d_rays, d_spheres, d_tree, d_hit_counts):
    """cuda/trace_sph.cuh:58-79.  ValueError unless len(rays) % 32 == 0."""
    _need(d_hit_counts, torch.int32, "d_hit_counts")
    if d_hit_counts.numel() < d_rays.shape[0]:
        raise ValueError("d_hit_counts holds fewer elements than there are rays")
    _check(_t_counts(*_trace_args(d_rays, d_spheres, d_tree), _dp(d_hit_counts), _stream()))


def trace_cumulative_sph(d_rays, d_spheres, d_tree, d_cumulated):
    """cuda/trace_sph.cuh:82-109."""
    _need(d_cumulated, torch.float32, "d_cumulated")
    if d_cumulated.numel() < d_rays.shape[0]:
        raise ValueError("d_cumulated holds fewer elements than there are rays")
    _check(_t_cum(*_trace_args(d_rays, d_spheres, d_tree), _dp(d_cumulated), _stream()))


def trace_stats_sph(d_rays, d_spheres, d_tree):
    """Traversal counters of the packet algorithm: dict(node_visits, leaf_visits,
    prims_staged, hits) summed over all packets (for the algorithmic-bytes figure)."""
    out = (_c.c_longlong * 4)()
    _check(_t_stats(*_trace_args(d_rays, d_spheres, d_tree), ctypes.byref(out), _stream()))
    return dict(node_visits=out[0], leaf_visits=out[1], prims_staged=out[2], hits=out[3])


def trace_packet_profile_sph(d_rays, d_spheres, d_tree):
    """Diagnostic: work counters of the production packet kernel."""
    import numpy as np
    npk = d_rays.shape[0] // 32
    out = np.zeros(4 + 4 * npk, np.int64)
    cnt = torch.empty(d_rays.shape[0], dtype=torch.int32, device=d_rays.device)
    fn = _sig("grace_b200_trace_packet_profile_f4", [_P, _P, _sz, _P, _sz, _TS, _P, _P, _c.c_int, _P])
    _check(fn(*_trace_args(d_rays, d_spheres, d_tree), _dp(cnt), out.ctypes.data_as(_P), 1, _stream()))
    return dict(node_steps=int(out[0]), leaf_visits=int(out[1]), prims_staged=int(out[2]),
                prims_kept=int(out[3]), per_packet=out[4:].reshape(npk, 4))


def trace_ray_cost_sph(d_rays, d_spheres, d_tree):
    """Diagnostic: (sphere tests, inner-node steps) per ray of the per-ray traversal."""
    n = d_rays.shape[0]
    tests = torch.empty(n, dtype=torch.int32, device=d_rays.device)
    steps = torch.empty(n, dtype=torch.int32, device=d_rays.device)
    fn = _sig("grace_b200_trace_ray_cost_f4", [_P, _P, _sz, _P, _sz, _TS, _P, _P, _P])
    _check(fn(*_trace_args(d_rays, d_spheres, d_tree), _dp(tests), _dp(steps), _stream()))
    return tests, steps


def _trace_lists(d_rays, d_spheres, d_tree, d_ray_offsets, sentinels):
    _need(d_ray_offsets, torch.int32, "d_ray_offsets")
    args = _trace_args(d_rays, d_spheres, d_tree)
    total = _c.c_longlong(0)
    _check(_t_hcount(*args, 1 if sentinels is not None else 0, _dp(d_ray_offsets),
                     ctypes.byref(total), _stream()))
    dev = d_rays.device
    if sentinels is None:
        idx = torch.empty(total.value, dtype=torch.int32, device=dev)
        integ = torch.empty(total.value, dtype=torch.float32, device=dev)
        dist = torch.empty(total.value, dtype=torch.float32, device=dev)
    else:
        idx = torch.full((total.value,), int(sentinels[0]), dtype=torch.int32, device=dev)
        integ = torch.full((total.value,), float(sentinels[1]), dtype=torch.float32, device=dev)
        dist = torch.full((total.value,), float(sentinels[2]), dtype=torch.float32, device=dev)
    if total.value > 0:
        _check(_t_hfill(*args, _dp(d_ray_offsets), _dp(idx), _dp(integ), _dp(dist), _stream()))
    return idx, integ, dist


def trace_sph(d_rays, d_spheres, d_tree, d_ray_offsets):
    """cuda/trace_sph.cuh:112-168.  Fills d_ray_offsets (exclusive scan of hit counts) and
    returns the callee-sized outputs (d_hit_indices, d_hit_integrals, d_hit_distances)."""
    return _trace_lists(d_rays, d_spheres, d_tree, d_ray_offsets, None)


def trace_with_sentinels_sph(d_rays, d_spheres, d_tree, d_ray_offsets, index_sentinel,
                             integral_sentinel, distance_sentinel):
    """cuda/trace_sph.cuh:171-241: every ray segment ends with one sentinel slot."""
    return _trace_lists(d_rays, d_spheres, d_tree, d_ray_offsets,
                        (index_sentinel, integral_sentinel, distance_sentinel))


class _DevArray:
    """A library-owned device buffer as torch sees it (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


_TILE_FN = _c.CFUNCTYPE(_c.c_int, _P, _sz, _sz, _P, _c.c_longlong, _P, _P, _P, _P)
_t_tiles = _sig("grace_b200_trace_sorted_tiles_f4",
                [_P, _P, _sz, _P, _sz, _TS, _sz, _sz, _TILE_FN, _P, _c.POINTER(_c.c_longlong), _P])


def trace_sorted_tiles(d_rays, d_spheres, d_tree, hit_budget, consume, rays_per_tile=0):
    """Sorted hit lists of a ray set too large for one trace_sph call (BASELINE config 4), streamed in
    ray tiles: count -> scan -> fill -> sort_by_distance -> consume(first_ray, offsets, indices,
    integrals, distances) per tile, in library-owned buffers of `hit_budget` hits that are reused for
    every tile.  The tensors passed to `consume` are views of those buffers, valid only during the
    call; torch work issued inside it runs on the library's second stream (the current stream of the
    callback), overlapping the counting traversal of the next tile.  Returns the total number of hits."""
    failure = []

    def tramp(user, first, n_rays, p_off, n_hits, p_idx, p_int, p_dist, stream):
        try:
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                off = torch.as_tensor(_DevArray(p_off, n_rays, "<i4"), device=d_rays.device)
                if n_hits > 0:
                    idx = torch.as_tensor(_DevArray(p_idx, n_hits, "<i4"), device=d_rays.device)
                    integ = torch.as_tensor(_DevArray(p_int, n_hits, "<f4"), device=d_rays.device)
                    dist = torch.as_tensor(_DevArray(p_dist, n_hits, "<f4"), device=d_rays.device)
                else:
                    idx = torch.empty(0, dtype=torch.int32, device=d_rays.device)
                    integ = dist = torch.empty(0, dtype=torch.float32, device=d_rays.device)
                consume(int(first), off, idx, integ, dist)
            return 0
        except BaseException as e:      # never let an exception cross the C frame
            failure.append(e)
            return 7

    cb = _TILE_FN(tramp)
    total = _c.c_longlong(0)
    rc = _t_tiles(*_trace_args(d_rays, d_spheres, d_tree), int(hit_budget), int(rays_per_tile), cb, None,
                  ctypes.byref(total), _stream())
    if failure:
        raise failure[0]
    _check(rc)
    return total.value


def sort_by_distance(d_hit_distances, d_ray_offsets, d_hit_indices, d_hit_data):
    """cuda/sort.cuh:100-131: per-ray stable sort by distance, in place."""
    _need(d_hit_distances, torch.float32, "d_hit_distances")
    _need(d_ray_offsets, torch.int32, "d_ray_offsets")
    _need(d_hit_indices, torch.int32, "d_hit_indices")
    if d_hit_data.element_size() != 4 or d_hit_data.numel() != d_hit_distances.numel():
        raise TypeError("d_hit_data must hold one 32-bit value per hit")
    _check(_sort_dist(context(), _dp(d_hit_distances), _dp(d_ray_offsets), d_ray_offsets.numel(),
                      d_hit_distances.numel(), _dp(d_hit_indices), _dp(d_hit_data), _stream()))


def exclusive_segmented_scan(d_segment_offsets, d_data, d_results):
    """cuda/scan.cuh:15-38.  d_data and d_results may be the same tensor."""
    _need(d_segment_offsets, torch.int32, "d_segment_offsets")
    _need(d_data, torch.float32, "d_data")
    _need(d_results, torch.float32, "d_results")
    if d_results.numel() != d_data.numel():
        raise ValueError("d_results must have one value per datum")
    _check(_segscan(context(), _dp(d_segment_offsets), d_segment_offsets.numel(), _dp(d_data), d_data.numel(),
                    _dp(d_results), _stream()))
    return d_results


def weighted_exclusive_segmented_scan(d_to_sum, d_weights, d_weight_map, d_segment_offsets, d_sum):
    """cuda/scan.cuh:45-58: scans d_weights[d_weight_map[i]] * d_to_sum[i]; d_weight_map is
    uint32 in the reference (int32 tensors hold the same bits)."""
    _need(d_to_sum, torch.float32, "d_to_sum")
    _need(d_weights, torch.float32, "d_weights")
    _need(d_weight_map, torch.int32, "d_weight_map")
    _need(d_segment_offsets, torch.int32, "d_segment_offsets")
    _need(d_sum, torch.float32, "d_sum")
    if d_weight_map.numel() != d_to_sum.numel() or d_sum.numel() != d_to_sum.numel():
        raise ValueError("d_weight_map and d_sum must have one value per datum")
    _check(_wsegscan(context(), _dp(d_to_sum), _dp(d_weights), _dp(d_weight_map), _dp(d_segment_offsets),
                     d_segment_offsets.numel(), d_to_sum.numel(), _dp(d_sum), _stream()))
    return d_sum


def offsets_to_segments(d_offsets, d_segments):
    """cuda/sort.cuh:20-41."""
    _need(d_offsets, torch.int32, "d_offsets")
    _need(d_segments, torch.int32, "d_segments")
    _check(_off2seg(context(), _dp(d_offsets), d_offsets.numel(), _dp(d_segments), d_segments.numel(), _stream()))
    return d_segments


def exclusive_scan(d_in, d_out=None):
    """thrust::exclusive_scan on int32 (trace_sph.cuh:135); returns (out, total tensor int64[1])."""
    _need(d_in, torch.int32, "d_in")
    d_out = d_in if d_out is None else d_out
    total = torch.zeros(1, dtype=torch.int64, device=d_in.device)
    _check(_scan(context(), _dp(d_in), _dp(d_out), d_in.numel(), _dp(total), _stream()))
    return d_out, total


# ----------------------------------------------------------------------------- rays
def _rays_out(d_rays, n):
    if d_rays is None:
        d_rays = torch.empty((n, 7), dtype=torch.float32, device="cuda")
    _need(d_rays, torch.float32, "d_rays", 7)
    if d_rays.shape[0] < n:
        raise ValueError("d_rays too small")   # the reference resizes; tensors cannot
    return d_rays


def uniform_random_rays(d_rays, ox, oy, oz, length, seed=1234):
    """cuda/gen_rays.cuh:26-51."""
    _need(d_rays, torch.float32, "d_rays", 7)
    _check(_rays_uniform(context(), _dp(d_rays), d_rays.shape[0], ox, oy, oz, length, -1, seed, _stream()))
    return d_rays


def uniform_random_rays_single_octant(d_rays, ox, oy, oz, length, octant=Octants.PPP, seed=1234):
    """cuda/gen_rays.cuh:60-96."""
    _need(d_rays, torch.float32, "d_rays", 7)
    _check(_rays_uniform(context(), _dp(d_rays), d_rays.shape[0], ox, oy, oz, length, int(octant), seed,
                         _stream()))
    return d_rays


def one_to_many_rays(d_rays, ox, oy, oz, d_points, sort_type=RaySortType.DirectionSort,
                     AABB_bot=None, AABB_top=None):
    """cuda/gen_rays.cuh:98-186.  d_points float32 [n, 3 or 4]."""
    if sort_type not in (0, 1, 2):
        raise ValueError("Ray sort type not recognized")     # gen_rays.cuh:124-130
    n = d_points.shape[0]
    d_rays = _rays_out(d_rays, n)
    hb = _h3(AABB_bot) if AABB_bot is not None else None
    ht = _h3(AABB_top) if AABB_top is not None else None
    _check(_rays_o2m(context(), _dp(d_rays), n, ox, oy, oz, _dp(d_points), d_points.shape[1],
                     int(sort_type), hb, ht, _stream()))
    return d_rays


def plane_parallel_random_rays(d_rays, width, height, base, w, h, length, seed=1234):
    """cuda/gen_rays.cuh:188-238."""
    d_rays = _rays_out(d_rays, width * height)
    _check(_rays_pp(context(), _dp(d_rays), width, height, _h3(base), _h3(w), _h3(h), length, seed,
                    _stream()))
    return d_rays


def orthographic_projection_rays(d_rays, resolution_x, resolution_y, camera_position, look_at,
                                 view_up, vertical_extent, length):
    """cuda/gen_rays.cuh:240-290."""
    d_rays = _rays_out(d_rays, resolution_x * resolution_y)
    _check(_rays_ortho(context(), _dp(d_rays), resolution_x, resolution_y, _h3(camera_position),
                       _h3(look_at), _h3(view_up), vertical_extent, length, _stream()))
    return d_rays


def pinhole_camera_rays(d_rays, resolution_x, resolution_y, camera_position, look_at, view_up,
                        FOVy, length):
    """cuda/gen_rays.cuh:292-399."""
    d_rays = _rays_out(d_rays, resolution_x * resolution_y)
    _check(_rays_pinhole(context(), _dp(d_rays), resolution_x, resolution_y, _h3(camera_position),
                         _h3(look_at), _h3(view_up), FOVy, length, _stream()))
    return d_rays


def healpix_rays(d_rays, nside, first_pixel, n_rays, ox, oy, oz, length):
    """HEALPix NESTED pixel-centre rays (RayVectorGeneration/src/chealpix/chealpix.c:459-467)."""
    d_rays = _rays_out(d_rays, n_rays)
    _check(_rays_healpix(context(), _dp(d_rays), n_rays, nside, first_pixel, ox, oy, oz, length,
                         _stream()))
    return d_rays


def gadget_info(path):
    """Header of a Gadget-2 type-1 file: (npart[6], mass[6], n_gas)."""
    np6 = (_c.c_longlong * 6)()
    m6 = (_c.c_double * 6)()
    ng = _c.c_longlong(0)
    _check(_gadget_info(os.fsencode(path), np6, m6, _c.byref(ng)))
    return list(np6), list(m6), int(ng.value)


def read_gadget(path, d_spheres=None):
    """tests/helper/read_gadget.cuh:161-167: gas positions + smoothing lengths as float4 records
    on the current device.  Asynchronous on the current stream after the last file read."""
    n = gadget_info(path)[2]
    if d_spheres is None:
        d_spheres = torch.empty((max(n, 1), 4), dtype=torch.float32, device="cuda")
    _need(d_spheres, torch.float32, "d_spheres", 4)
    got = _c.c_size_t(0)
    _check(_gadget_read(context(), os.fsencode(path), _dp(d_spheres), d_spheres.shape[0], _c.byref(got), _stream()))
    return d_spheres[: got.value]


def write_gadget(path, spheres, n_other=0, other_has_mass_block=False):
    """Driver utility: host float4 records -> a Gadget-2 type-1 file the reference's reader loads."""
    h = spheres.detach().cpu().contiguous() if isinstance(spheres, torch.Tensor) else torch.from_numpy(spheres).contiguous()
    if h.dtype != torch.float32 or h.dim() != 2 or h.shape[1] != 4:
        raise TypeError("spheres must be float32 [N, 4]")
    _check(_gadget_write(os.fsencode(path), _c.c_void_p(h.data_ptr()), h.shape[0], n_other, int(other_has_mass_block)))


def synth_gadget_spheres(n, seed=1234, device=None):
    """Synthetic Gadget-shaped snapshot, generated on the device."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    s = torch.empty((n, 4), dtype=torch.float32, device=dev)
    _check(_synth(context(), _dp(s), n, seed, _stream()))
    return s
