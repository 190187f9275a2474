"""Bit-level parity with the REFERENCE'S OWN CUDA IMPLEMENTATION run on the same GPU
(oracle/_ref/ref_driver = GRACE's headers, patched only for CUDA-12 API removals, called
through GRACE's public API).  Three-way: reference CUDA == CPU oracle == this repo's CUDA."""
import numpy as np
import pytest
import torch

import refrun
from util import clustered_spheres, uniform_spheres, isotropic_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def have_ref():
    if not refrun.available():
        pytest.skip("oracle/_ref/ref_driver not built")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def ours(gb, s, rays, mpl, bits):
    d_s = dev(s)
    (gb.morton_keys30_sort_sph if bits == 30 else gb.morton_keys63_sort_sph)(d_s)
    deltas = torch.empty(len(s) + 1, dtype=torch.float32, device="cuda")
    gb.euclidean_deltas_sph(d_s, deltas)
    tree = gb.Tree(len(s), mpl)
    gb.ALBVH_sph(d_s, deltas, tree)
    d_r = dev(rays)
    cnt = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    cum = torch.empty(len(rays), dtype=torch.float32, device="cuda")
    gb.trace_hitcounts_sph(d_r, d_s, tree, cnt)
    gb.trace_cumulative_sph(d_r, d_s, tree, cum)
    off = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    idx, integ, dist = gb.trace_sph(d_r, d_s, tree, off)
    gb.sort_by_distance(dist, off, idx, integ)
    return dict(spheres_sorted=host(d_s), deltas=host(deltas), leaves=host(tree.leaves), nodes=host(tree.nodes),
                root=int(tree.root_index_ptr.item()), hitcounts=host(cnt), cumulative=host(cum),
                offsets=host(off), hit_idx=host(idx), hit_integral=host(integ), hit_dist=host(dist))


def assert_same_hits(ref, got):
    """Per-ray hit lists sorted by distance: distances and integrals bit-exact in order;
    indices equal except inside groups of exactly tied distances, where the reference's
    order is its emission order and ours is ascending index (both are 'stable')."""
    assert np.array_equal(ref["offsets"], got["offsets"])
    assert np.array_equal(ref["hit_dist"].view(np.uint32), got["hit_dist"].view(np.uint32))
    same = ref["hit_idx"] == got["hit_idx"]
    if not same.all():
        d = ref["hit_dist"]
        bad = np.nonzero(~same)[0]
        tied = (d[bad] == d[np.maximum(bad - 1, 0)]) | (d[bad] == d[np.minimum(bad + 1, len(d) - 1)])
        assert tied.all()
    ends = np.append(ref["offsets"][1:], len(ref["hit_idx"]))
    for b, e in list(zip(ref["offsets"], ends))[::97]:
        assert sorted(ref["hit_idx"][b:e]) == sorted(got["hit_idx"][b:e])
        assert np.array_equal(np.sort(ref["hit_integral"][b:e]), np.sort(got["hit_integral"][b:e]))


@pytest.mark.parametrize("bits,mpl,data", [(30, 32, "clustered"), (63, 32, "clustered"), (30, 8, "uniform"),
                                           (30, 1, "uniform")])
def test_three_way_parity(gb, orc, have_ref, bits, mpl, data):
    n = 1 << 16
    s = clustered_spheres(n, seed=11) if data == "clustered" else uniform_spheres(n, seed=12, rmax=0.03)
    rays = isotropic_rays(2048, seed=13)
    ref, info = refrun.run(s, rays, mpl, bits, iters=0, lists=True)
    got = ours(gb, s, rays, mpl, bits)
    # --- reference CUDA vs this repo's CUDA: bit-exact
    assert np.array_equal(ref["spheres_sorted"].view(np.uint32), got["spheres_sorted"].view(np.uint32))
    assert np.array_equal(ref["deltas"].view(np.uint32), got["deltas"].view(np.uint32))
    assert np.array_equal(ref["leaves"][:, :2], got["leaves"][:, :2])
    assert ref["root"] == got["root"]
    assert np.array_equal(ref["nodes"], got["nodes"])
    assert np.array_equal(ref["hitcounts"], got["hitcounts"])
    rel = np.abs(ref["cumulative"] - got["cumulative"]) / np.maximum(np.abs(ref["cumulative"]), 1e-30)
    assert rel.max() <= 1e-5
    assert np.array_equal(ref["cumulative"].view(np.uint32), got["cumulative"].view(np.uint32))
    assert_same_hits(ref, got)
    # --- reference CUDA vs CPU oracle: bit-exact (this is what pins the oracle)
    hs, _, _ = orc.sort_spheres(s, bits)
    assert np.array_equal(ref["spheres_sorted"].view(np.uint32), hs.view(np.uint32))
    htree = orc.build_tree(hs, orc.deltas_euclid(hs), mpl)
    assert np.array_equal(ref["nodes"], htree.nodes) and ref["root"] == htree.root
    assert np.array_equal(ref["leaves"][:, :2], htree.leaves[:, :2])
    assert np.array_equal(ref["hitcounts"], orc.trace_hitcounts(rays, hs, htree))
    assert np.array_equal(ref["cumulative"].view(np.uint32), orc.trace_cumulative(rays, hs, htree).view(np.uint32))


def test_uniform_random_rays_bit_exact(gb, have_ref):
    """Same device, same cuRAND sub-sequences, same double-precision normalisation, same
    direction keys and a stable sort: the generated rays must be identical bit for bit."""
    s = uniform_spheres(4096, seed=1)
    for n_rays, seed in ((32 * 1000, 1234), (1 << 17, 7)):
        ref, _ = refrun.run(s, "gen:%d:%d:0.5:0.25:0.125:2.0" % (n_rays, seed), 32, 30, iters=0, lists=False)
        rays = torch.empty((n_rays, 7), dtype=torch.float32, device="cuda")
        gb.uniform_random_rays(rays, 0.5, 0.25, 0.125, 2.0, seed)
        assert np.array_equal(ref["rays"].view(np.uint32), host(rays).view(np.uint32))


def test_uniform_random_rays_golden(gb):
    """Golden rays recorded from the reference on a 148-SM B200 (the generator's state count
    depends on the SM count by construction, cuda/kernels/gen_rays.cuh:428-438)."""
    import os
    if torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("golden rays were recorded on a 148-SM device")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "uniform_random_rays_4096_seed1234_b200.npz"))
    rays = torch.empty((4096, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays(rays, 0.5, 0.5, 0.5, 2.0, 1234)
    assert np.array_equal(g["rays"].view(np.uint32), host(rays).view(np.uint32))


# ----------------------------------------------------------------------------- ray generators
def gen_points(n, seed=21):
    return np.random.default_rng(seed).random((n, 3), dtype=np.float32)


def our_gens(gb, points, n_random, seed, rx, ry):
    """The product's generators with the parameters oracle/ref_gen_driver.cu hard-codes."""
    P = refrun.GEN_PARAMS
    out = {}
    r = torch.empty((n_random, 7), dtype=torch.float32, device="cuda")
    gb.uniform_random_rays_single_octant(r, *P["octant_origin"], P["octant_length"], gb.Octants.MPM, seed)
    out["octant"] = host(r)
    d_pts = dev(points)
    for name, st in (("o2m_nosort", gb.RaySortType.NoSort), ("o2m_dirsort", gb.RaySortType.DirectionSort)):
        out[name] = host(gb.one_to_many_rays(None, *P["o2m_origin"], d_pts, st))
    out["o2m_endsort_aabb"] = host(gb.one_to_many_rays(None, *P["o2m_origin"], d_pts, gb.RaySortType.EndPointSort,
                                                       P["o2m_aabb"][0], P["o2m_aabb"][1]))
    out["plane_parallel"] = host(gb.plane_parallel_random_rays(None, rx, ry, P["pp_base"], P["pp_w"], P["pp_h"],
                                                               P["pp_length"], seed))
    out["ortho"] = host(gb.orthographic_projection_rays(None, rx, ry, P["cam_pos"], P["look_at"], P["view_up"],
                                                        P["ortho_extent"], P["cam_length"]))
    out["pinhole"] = host(gb.pinhole_camera_rays(None, rx, ry, P["cam_pos"], P["look_at"], P["view_up"],
                                                 P["pinhole_fovy"], P["cam_length"]))
    return out


@pytest.mark.parametrize("n_pts,n_random,seed,rx,ry", [(5000, 32 * 300, 99, 96, 64), (1 << 16, 1 << 16, 1234, 257, 129)])
def test_all_ray_generators_bit_exact(gb, n_pts, n_random, seed, rx, ry):
    """Same device: cuRAND sub-sequences, double-precision normalisation, image-plane arithmetic,
    direction / end-point keys and the stable sort must give the reference's rays bit for bit."""
    if not refrun.gens_available():
        pytest.skip("oracle/_ref/ref_gen_driver not built")
    pts = gen_points(n_pts)
    ref = refrun.run_gens(pts, n_random, seed, rx, ry)
    got = our_gens(gb, pts, n_random, seed, rx, ry)
    for k in refrun.GEN_NAMES:
        assert ref[k].shape == got[k].shape, k
        assert np.array_equal(ref[k].view(np.uint32), got[k].view(np.uint32)), k


def test_deterministic_generators_golden(gb):
    """one_to_many / orthographic / pinhole rays do not depend on the device: golden rays recorded
    from the reference's CUDA build (tests/golden/make_golden.py) must be reproduced anywhere."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ray_generators_ref.npz"))
    got = our_gens(gb, g["points"], 64, 5, int(g["res_x"]), int(g["res_y"]))
    for k in ("o2m_nosort", "o2m_dirsort", "o2m_endsort_aabb", "ortho", "pinhole"):
        assert np.array_equal(g[k].view(np.uint32), got[k].view(np.uint32)), k
    if torch.cuda.get_device_properties(0).multi_processor_count == 148:   # cuRAND state count ~ SMs
        for k in ("octant", "plane_parallel"):
            got2 = our_gens(gb, g["points"], int(g["n_random"]), int(g["seed"]), int(g["res_x"]), int(g["res_y"]))
            assert np.array_equal(g[k].view(np.uint32), got2[k].view(np.uint32)), k
