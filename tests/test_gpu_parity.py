"""GPU parity tests: the CUDA path (through the C ABI, via the ctypes mirror of the
reference's API) against the CPU oracle on the same seeded inputs.
Bit-exact for keys, order, tree arrays, hit counts and hit lists; column densities
are also required bit-exact here (both sum hits in ascending primitive order),
with 1e-5 relative as the documented contract."""
import numpy as np
import pytest
import torch

from util import uniform_spheres, clustered_spheres, isotropic_rays, ortho_rays_z

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def as_u(t):
    a = host(t)
    return a.view(np.uint32) if a.dtype == np.int32 else a.view(np.uint64)


def build_both(gb, orc, s, mpl, bits=30, bot=None, top=None, delta="euclid"):
    """Sort + deltas + tree on the GPU and in the oracle from the same input."""
    d_s = dev(s)
    fn = gb.morton_keys30_sort_sph if bits == 30 else gb.morton_keys63_sort_sph
    d_keys = fn(d_s, bot, top, return_keys=True)
    hs, hk, perm = orc.sort_spheres(s, bits, bot, top)
    n = len(s)
    if delta == "euclid":
        d_d = torch.empty(n + 1, dtype=torch.float32, device="cuda")
        gb.euclidean_deltas_sph(d_s, d_d)
        hd = orc.deltas_euclid(hs)
    elif delta == "sarea":
        d_d = torch.empty(n + 1, dtype=torch.float32, device="cuda")
        gb.surface_area_deltas_sph(d_s, d_d)
        hd = orc.deltas_sarea(hs)
    else:
        d_d = torch.empty(n + 1, dtype=d_keys.dtype, device="cuda")
        gb.XOR_deltas_sph(d_keys, d_d)
        hd = orc.deltas_xor(hk)
    tree = gb.Tree(n, mpl)
    gb.ALBVH_sph(d_s, d_d, tree)
    htree = orc.build_tree(hs, hd, mpl)
    return d_s, d_keys, d_d, tree, hs, hk, hd, htree


def assert_tree_equal(tree, htree):
    assert tree.n_leaves == htree.n_leaves
    assert np.array_equal(host(tree.leaves)[:, :2], htree.leaves[:, :2])
    assert int(tree.root_index_ptr.item()) == htree.root
    assert np.array_equal(host(tree.nodes), htree.nodes)


# ----------------------------------------------------------------------------- keys
def test_bounds_and_minmax(gb, orc):
    s = clustered_spheres(100003, seed=1)
    d_s = dev(s)
    lo, hi = orc.bounds(s)
    assert np.array_equal(host(gb.min_vec3(d_s)), lo)
    assert np.array_equal(host(gb.max_vec3(d_s)), hi)
    assert np.array_equal(host(gb.max_vec4(d_s)), s.max(0))
    assert gb.min_max_x(d_s) == (float(s[:, 0].min()), float(s[:, 0].max()))


@pytest.mark.parametrize("bits", [30, 63])
@pytest.mark.parametrize("explicit", [True, False])
def test_morton_keys(gb, orc, bits, explicit):
    s = uniform_spheres(77777, seed=2) * np.float32(2.0) - np.float32(1.0)  # [-1, 1)
    d_s = dev(s)
    keys = torch.empty(len(s), dtype=torch.int32 if bits == 30 else torch.int64, device="cuda")
    if explicit:
        bot, top = np.full(3, -1, np.float32), np.full(3, 1, np.float32)
        gb.morton_keys_sph(d_s, keys, bot, top)
    else:
        bot, top = orc.bounds(s)
        gb.morton_keys_sph(d_s, keys)
    assert np.array_equal(as_u(keys), orc.morton_keys(s, bot, top, bits))


def test_morton_keys_out_of_range_and_nan(gb, orc):
    # cvt.rzi saturation / NaN -> 0 (F2I.TRUNC semantics of the reference kernel)
    s = np.array([[-5, 0.5, 2.0, 1], [np.nan, 1.0, 0.999, 1], [0.25, np.inf, -np.inf, 1],
                  [1e30, -1e30, 0.5, 1]], np.float32)
    s = np.tile(s, (40, 1))
    bot, top = np.zeros(3, np.float32), np.ones(3, np.float32)
    for bits, dt in ((30, torch.int32), (63, torch.int64)):
        keys = torch.empty(len(s), dtype=dt, device="cuda")
        gb.morton_keys_sph(dev(s), keys, bot, top)
        assert np.array_equal(as_u(keys), orc.morton_keys(s, bot, top, bits))


# ----------------------------------------------------------------------------- sort
@pytest.mark.parametrize("n", [1, 31, 6144, 6145, 200001])
@pytest.mark.parametrize("kbits", [32, 64])
def test_sort_pairs_stable(gb, orc, n, kbits):
    rng = np.random.default_rng(n + kbits)
    if kbits == 32:
        keys = rng.integers(0, 1 << 12, n).astype(np.uint32) * np.uint32(262147)  # many duplicates
        d_k = dev(keys.view(np.int32))
    else:
        keys = rng.integers(0, 1 << 16, n).astype(np.uint64) * np.uint64(281474976710677)
        d_k = dev(keys.view(np.int64))
    vals = rng.random((n, 4), dtype=np.float32)
    d_v = dev(vals)
    perm = gb.sort_by_key(d_k, d_v, return_perm=True)
    ref = orc.sort_perm(keys)
    assert np.array_equal(host(perm), ref)
    assert np.array_equal(as_u(d_k), keys[ref])
    assert np.array_equal(host(d_v), vals[ref])


def test_sort_ray_payload_and_key_bits(gb, orc):
    rng = np.random.default_rng(5)
    n = 50021
    keys = rng.integers(0, 1 << 30, n).astype(np.uint32)
    rays = rng.random((n, 7), dtype=np.float32)
    d_k, d_r = dev(keys.view(np.int32)), dev(rays)
    gb.sort_by_key(d_k, d_r, key_bits=30)
    ref = orc.sort_perm(keys)
    assert np.array_equal(host(d_r), rays[ref])
    vals = np.arange(n, dtype=np.int32)
    d_k, d_v = dev(keys.view(np.int32)), dev(vals)
    gb.sort_by_key(d_k, d_v)
    assert np.array_equal(host(d_v), ref)


@pytest.mark.parametrize("bits", [30, 63])
def test_morton_sort_spheres(gb, orc, bits):
    s = clustered_spheres(150000, seed=7)       # clustered => many duplicate 30-bit keys
    d_s = dev(s)
    fn = gb.morton_keys30_sort_sph if bits == 30 else gb.morton_keys63_sort_sph
    keys = fn(d_s, return_keys=True)
    hs, hk, perm = orc.sort_spheres(s, bits)
    assert np.array_equal(as_u(keys), hk)
    assert np.array_equal(host(d_s), hs)


# ----------------------------------------------------------------------------- deltas + tree
def test_deltas(gb, orc):
    s = clustered_spheres(60000, seed=8)
    hs, hk, _ = orc.sort_spheres(s, 30)
    d_s = dev(hs)
    d = torch.empty(len(s) + 1, dtype=torch.float32, device="cuda")
    gb.euclidean_deltas_sph(d_s, d)
    assert np.array_equal(host(d).view(np.uint32), orc.deltas_euclid(hs).view(np.uint32))
    gb.surface_area_deltas_sph(d_s, d)
    assert np.array_equal(host(d).view(np.uint32), orc.deltas_sarea(hs).view(np.uint32))
    for bits, dt in ((30, torch.int32), (63, torch.int64)):
        hk = orc.morton_keys(hs, *orc.bounds(hs), bits)
        dk = dev(hk.view(np.int32 if bits == 30 else np.int64))
        dd = torch.empty(len(s) + 1, dtype=dt, device="cuda")
        gb.XOR_deltas_sph(dk, dd)
        assert np.array_equal(as_u(dd), orc.deltas_xor(hk))


@pytest.mark.parametrize("mpl", [1, 2, 8, 32, 100])
@pytest.mark.parametrize("data", ["uniform", "clustered"])
def test_tree_bit_exact(gb, orc, mpl, data):
    n = 120000
    s = uniform_spheres(n, seed=mpl) if data == "uniform" else clustered_spheres(n, seed=mpl)
    out = build_both(gb, orc, s, mpl)
    assert_tree_equal(out[3], out[7])


@pytest.mark.parametrize("delta,bits", [("xor", 30), ("xor", 63), ("sarea", 30)])
def test_tree_other_deltas(gb, orc, delta, bits):
    s = clustered_spheres(90000, seed=21)
    out = build_both(gb, orc, s, 16, bits=bits, delta=delta)
    assert_tree_equal(out[3], out[7])


@pytest.mark.parametrize("n", [34, 63, 64, 65, 97, 8127, 8128, 8129, 8160, 8161, 16257, 100003])
def test_tree_default_leaf_size_edges(gb, orc, n):
    """max_per_leaf = 32 takes the sliding-window kernel (one thread per block of 32 deltas, a CTA emits 254
    blocks = 8128 nodes): sizes around its block and tile boundaries, and coarse coordinates so that many
    deltas are EQUAL (the leftmost-maximum tie rule decides where leaves end)."""
    s = clustered_spheres(n, seed=n)
    out = build_both(gb, orc, s, 32)
    assert_tree_equal(out[3], out[7])
    q = s.copy()
    q[:, :3] = np.round(q[:, :3] * 64.0) / 64.0          # a 64^3 lattice: runs of equal keys and equal distances
    out = build_both(gb, orc, q, 32)
    assert_tree_equal(out[3], out[7])


@pytest.mark.parametrize("delta,bits", [("xor", 30), ("xor", 63), ("sarea", 30)])
def test_tree_default_leaf_size_other_deltas(gb, orc, delta, bits):
    s = clustered_spheres(70001, seed=5)
    s[:, :3] = np.round(s[:, :3] * 256.0) / 256.0
    out = build_both(gb, orc, s, 32, bits=bits, delta=delta)
    assert_tree_equal(out[3], out[7])


def test_tree_tiny_and_errors(gb, orc):
    s = np.array([[-0.5, -0.5, -0.5, 0.2], [0.5, 0.5, 0.5, 0.2]], np.float32)
    out = build_both(gb, orc, s, 1, bot=-np.ones(3, np.float32), top=np.ones(3, np.float32))
    assert_tree_equal(out[3], out[7])
    with pytest.raises(ValueError):           # albvh.cuh:795-799
        d = torch.empty(3, dtype=torch.float32, device="cuda")
        gb.ALBVH_sph(dev(s), d, gb.Tree(2, 2))
    s3 = uniform_spheres(33, seed=1)
    out = build_both(gb, orc, s3, 32)
    assert_tree_equal(out[3], out[7])


# ----------------------------------------------------------------------------- trace
@pytest.fixture(params=["packet", "ray", "packet_ref"], autouse=True)
def trace_mode(request, gb):
    """Every test below runs under both traversal schedules."""
    gb.set_trace_mode(request.param)
    yield request.param
    gb.set_trace_mode("packet")


@pytest.fixture(scope="module")
def scene(gb, orc):
    s = clustered_spheres(1 << 17, seed=3)
    d_s, _, _, tree, hs, _, _, htree = build_both(gb, orc, s, 32)
    rays = isotropic_rays(1 << 12, origin=(0.5, 0.5, 0.5), length=2.0, seed=4)
    return d_s, tree, hs, htree, rays


def test_hitcounts_exact(gb, orc, scene):
    d_s, tree, hs, htree, rays = scene
    out = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    gb.trace_hitcounts_sph(dev(rays), d_s, tree, out)
    ref = orc.trace_hitcounts(rays, hs, htree)
    assert np.array_equal(host(out), ref)
    sub = slice(0, 256)
    assert np.array_equal(host(out)[sub], orc.brute_hitcounts(rays[sub], hs))


def test_traversal_counters_match_oracle(gb, orc, scene):
    # the algorithmic-bytes figure of bench.py rests on these counters
    d_s, tree, hs, htree, rays = scene
    st = gb.trace_stats_sph(dev(rays), d_s, tree)
    counts, stats, depth = orc.trace(rays, hs, htree, 0, with_stats=True)
    assert st["node_visits"] == int(stats[:, 0].sum())
    assert st["leaf_visits"] == int(stats[:, 1].sum())
    assert st["prims_staged"] == int(stats[:, 2].sum())
    assert st["hits"] == int(counts.sum())
    assert depth <= 64          # reference STACK_SIZE (kernel_config.h:13)


def test_hitcounts_config1(gb, orc):
    # BASELINE config 1: 2^16 uniform spheres (radii < 0.1), 2^14 isotropic rays from the
    # box centre, length 2, explicit bounds (0,0,0)-(1,1,1); checked against brute force.
    s = uniform_spheres(1 << 16, seed=1234)
    bot, top = np.zeros(3, np.float32), np.ones(3, np.float32)
    d_s, _, _, tree, hs, _, _, htree = build_both(gb, orc, s, 32, bot=bot, top=top)
    rays = isotropic_rays(1 << 14, seed=1234)
    out = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    gb.trace_hitcounts_sph(dev(rays), d_s, tree, out)
    assert np.array_equal(host(out), orc.brute_hitcounts(rays, hs))


def test_cumulative(gb, orc, scene):
    d_s, tree, hs, htree, rays = scene
    out = torch.empty(len(rays), dtype=torch.float32, device="cuda")
    gb.trace_cumulative_sph(dev(rays), d_s, tree, out)
    ref = orc.trace_cumulative(rays, hs, htree)
    got = host(out)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
    assert rel.max() <= 1e-5          # the contract (north_star)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))   # and in fact bit-exact


def test_hit_lists_and_sort(gb, orc, scene):
    d_s, tree, hs, htree, rays = scene
    rays = rays[:1024]
    off = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    idx, integ, dist = gb.trace_sph(dev(rays), d_s, tree, off)
    roff, ridx, rinteg, rdist = orc.trace_hits(rays, hs, htree)
    assert np.array_equal(host(off), roff)
    assert np.array_equal(host(idx), ridx)
    assert np.array_equal(host(dist).view(np.uint32), rdist.view(np.uint32))
    assert np.array_equal(host(integ).view(np.uint32), rinteg.view(np.uint32))
    gb.sort_by_distance(dist, off, idx, integ)
    sd, si, sg = orc.sort_by_distance(rdist, roff, ridx, rinteg)
    assert np.array_equal(host(dist).view(np.uint32), sd.view(np.uint32))
    assert np.array_equal(host(idx), si)
    assert np.array_equal(host(integ).view(np.uint32), sg.view(np.uint32))


@pytest.mark.parametrize("pool", ["auto", "tiny", "none"])
@pytest.mark.parametrize("budget", [8, 100, 1000])
def test_packet_splitting_is_exact(gb, orc, scene, budget, pool, trace_mode):
    """Over-budget units are suspended and their pending subtrees handed to tasks (hit counts,
    column densities: terms recorded in chunk chains and folded in traversal order) or resumed as
    ray-subset tasks (hit lists); with a tiny budget nearly every packet goes through every
    round, nested.  With a tiny or empty term pool most tasks abort and the fold launch walks
    their subtrees itself.  Results must not change."""
    if trace_mode != "packet":
        pytest.skip("splitting exists only in the production packet schedule")
    d_s, tree, hs, htree, rays = scene
    gb.set_trace_budget(budget, eager=True)
    gb.set_trace_pool({"auto": 0, "tiny": 20 * 16384, "none": 1}[pool])
    try:
        d_rays = dev(rays)
        cnt = torch.empty(len(rays), dtype=torch.int32, device="cuda")
        cum = torch.empty(len(rays), dtype=torch.float32, device="cuda")
        gb.trace_hitcounts_sph(d_rays, d_s, tree, cnt)
        gb.trace_cumulative_sph(d_rays, d_s, tree, cum)
        assert gb.device_error() == 0
        assert np.array_equal(host(cnt), orc.trace_hitcounts(rays, hs, htree))
        ref = orc.trace_cumulative(rays, hs, htree)
        assert np.array_equal(host(cum).view(np.uint32), ref.view(np.uint32))
        sub = rays[:1024]
        off = torch.empty(len(sub), dtype=torch.int32, device="cuda")
        idx, integ, dist = gb.trace_sph(dev(sub), d_s, tree, off)
        roff, ridx, rinteg, rdist = orc.trace_hits(sub, hs, htree)
        assert np.array_equal(host(off), roff) and np.array_equal(host(idx), ridx)
        assert np.array_equal(host(integ).view(np.uint32), rinteg.view(np.uint32))
        assert np.array_equal(host(dist).view(np.uint32), rdist.view(np.uint32))
    finally:
        gb.set_trace_budget(1024)
        gb.set_trace_pool(0)


def test_hit_lists_one_traversal_equals_two(gb, orc, scene, trace_mode):
    """trace_sph records the hits during the counting traversal (work stealing, nested thefts) and copies them;
    the lists must equal those of count-then-fill, entry for entry, also when the recording pool overflows
    (fallback to the second traversal) -- after which the context asks for a pool that fits."""
    if trace_mode != "packet":
        pytest.skip("recording exists only in the production packet schedule")
    d_s, tree, hs, htree, rays = scene
    sub = dev(rays)
    nr = len(rays)
    off2 = torch.empty(nr, dtype=torch.int32, device="cuda")
    gb.set_hit_list_passes(2)
    try:
        want = gb.trace_sph(sub, d_s, tree, off2)
    finally:
        gb.set_hit_list_passes(1)
    for budget, eager in ((8, True), (16, False)):
        gb.set_trace_budget(budget, eager=eager)
        try:
            off = torch.empty(nr, dtype=torch.int32, device="cuda")
            got = gb.trace_sph(sub, d_s, tree, off)
            st = gb.trace_balance_stats()
            assert st["overflow"] == 0, "the default pool must hold the hits of these rays"
            if eager:
                assert st["tasks"] > 0, "an eager budget of 8 steps must lead to thefts"
            assert torch.equal(off, off2)
            for a, b in zip(got, want):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32))
        finally:
            gb.set_trace_budget(1024)
    # a pool of 64 chunks overflows: same lists through the fallback, and the next call's pool is sized from this one
    gb.set_trace_pool(64 * 8192)
    try:
        for expect_overflow in (1, 0):
            off = torch.empty(nr, dtype=torch.int32, device="cuda")
            got = gb.trace_sph(sub, d_s, tree, off)
            assert gb.trace_balance_stats()["overflow"] == expect_overflow
            assert torch.equal(off, off2)
            for a, b in zip(got, want):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    finally:
        gb.set_trace_pool(0)
    # with sentinels: every ray's segment ends with one slot holding them
    idx, integ, dist = gb.trace_with_sentinels_sph(sub, d_s, tree, off, -7, -1.0, 1e30)
    o = off.cpu().numpy().astype(np.int64); o2 = off2.cpu().numpy().astype(np.int64)
    assert np.array_equal(o, o2 + np.arange(nr))
    ends = np.concatenate([o[1:], [idx.numel()]]) - 1
    assert np.all(idx.cpu().numpy()[ends] == -7) and np.all(dist.cpu().numpy()[ends] == np.float32(1e30))
    keep = np.ones(idx.numel(), bool); keep[ends] = False
    assert np.array_equal(idx.cpu().numpy()[keep], want[0].cpu().numpy())
    assert np.array_equal(dist.cpu().numpy()[keep].view(np.uint32), want[2].cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("mpl", [1, 8, 64, 100])
def test_hit_lists_other_leaf_sizes(gb, orc, mpl, trace_mode):
    """Hit lists recorded during the counting traversal for leaves of 1 ... 100 primitives (the 32-, 64- and
    128-sphere staging variants of the packet kernel), thefts forced: equal to the oracle entry for entry, and
    to the two-traversal scheme."""
    if trace_mode != "packet":
        pytest.skip("recording exists only in the production packet schedule")
    s = clustered_spheres(40000, seed=100 + mpl)
    d_s, _, _, tree, hs, _, _, htree = build_both(gb, orc, s, mpl)
    rays = isotropic_rays(1024, origin=(0.5, 0.5, 0.5), length=2.0, seed=mpl)
    roff, ridx, rinteg, rdist = orc.trace_hits(rays, hs, htree)
    gb.set_trace_budget(8, eager=True)
    try:
        for passes in (1, 2):
            gb.set_hit_list_passes(passes)
            off = torch.empty(len(rays), dtype=torch.int32, device="cuda")
            idx, integ, dist = gb.trace_sph(dev(rays), d_s, tree, off)
            assert gb.device_error() == 0
            assert np.array_equal(host(off), roff) and np.array_equal(host(idx), ridx)
            assert np.array_equal(host(integ).view(np.uint32), rinteg.view(np.uint32))
            assert np.array_equal(host(dist).view(np.uint32), rdist.view(np.uint32))
    finally:
        gb.set_hit_list_passes(1)
        gb.set_trace_budget(1024)


def test_default_splitting_small_launches(gb, orc, scene):
    """The default policy on launches far smaller than the grid (every packet is suspended at the
    budget and its subtrees spread over the idle warps): results still exact."""
    d_s, tree, hs, htree, rays = scene
    for n in (64, 256, 4096):
        sub = np.ascontiguousarray(rays[:n])
        cnt = torch.empty(n, dtype=torch.int32, device="cuda")
        cum = torch.empty(n, dtype=torch.float32, device="cuda")
        gb.trace_hitcounts_sph(dev(sub), d_s, tree, cnt)
        gb.trace_cumulative_sph(dev(sub), d_s, tree, cum)
        assert gb.device_error() == 0
        assert np.array_equal(host(cnt), orc.trace_hitcounts(sub, hs, htree))
        ref = orc.trace_cumulative(sub, hs, htree)
        assert np.array_equal(host(cum).view(np.uint32), ref.view(np.uint32))


def test_axis_aligned_and_degenerate_directions(gb, orc, scene):
    """Rays with zero direction components (orthographic projections along an axis, -0.0, rays
    inside a box face plane): 1/d is infinite and the slab arithmetic must not turn into
    inf - inf.  Counts and column densities equal brute force in every schedule."""
    from util import ortho_rays_z
    d_s, tree, hs, htree, _ = scene
    rays = ortho_rays_z(32, -0.05, 1.05)                       # along -z, origins differ
    extra = rays[:96].copy()
    extra[:32, 0] = -0.0                                        # negative zero component
    extra[32:64, :3] = (1.0, 0.0, 0.0); extra[32:64, 3:6] = (-0.5, 0.5, 0.5)   # along +x, common origin
    extra[64:96, :3] = (0.0, -1.0, 0.0); extra[64:96, 4] = 1.5; extra[64:96, 5] = hs[1234, 2]
    rays = np.ascontiguousarray(np.concatenate([rays, extra]))
    d_rays = dev(rays)
    cnt = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    cum = torch.empty(len(rays), dtype=torch.float32, device="cuda")
    gb.trace_hitcounts_sph(d_rays, d_s, tree, cnt)
    gb.trace_cumulative_sph(d_rays, d_s, tree, cum)
    assert gb.device_error() == 0
    ref = orc.brute_hitcounts(rays, hs)
    assert ref.sum() > 0
    assert np.array_equal(host(cnt), ref)
    assert np.array_equal(host(cum).view(np.uint32), orc.brute_cumulative(rays, hs).view(np.uint32))


def test_trace_empty_and_degenerate_inputs(gb, orc, scene):
    """No rays is a no-op; zero-length rays hit nothing; rays with NaN/inf components must not
    hang or fault and must not disturb their packet neighbours.  (What a NaN ray itself "hits" is
    not a contract: every comparison of sphere_hit is false for NaN, so it accepts whatever it is
    tested against, generic/intersect.h:38-48, and that depends on the schedule.)"""
    d_s, tree, hs, htree, rays = scene
    empty = torch.empty((0, 7), dtype=torch.float32, device="cuda")
    gb.trace_hitcounts_sph(empty, d_s, tree, torch.empty(0, dtype=torch.int32, device="cuda"))
    gb.trace_cumulative_sph(empty, d_s, tree, torch.empty(0, dtype=torch.float32, device="cuda"))
    off = torch.empty(0, dtype=torch.int32, device="cuda")
    idx, integ, dist = gb.trace_sph(empty, d_s, tree, off)
    assert idx.numel() == 0
    r = rays[:64].copy()
    r[0:8, 6] = 0.0                       # zero length
    r[8, 0] = np.nan                      # NaN direction component
    r[9, 3] = np.nan                      # NaN origin
    r[10, 6] = np.inf                     # infinite length
    r[11, :3] = 0.0                       # null direction
    d_r = dev(r)
    cnt = torch.empty(64, dtype=torch.int32, device="cuda")
    cum = torch.empty(64, dtype=torch.float32, device="cuda")
    gb.trace_hitcounts_sph(d_r, d_s, tree, cnt)
    gb.trace_cumulative_sph(d_r, d_s, tree, cum)
    torch.cuda.synchronize()
    assert gb.device_error() == 0
    c = host(cnt)
    assert (c[0:8] == 0).all()
    ok = np.ones(64, bool)
    ok[8:12] = False                      # the oracle's own treatment of NaN/inf rays is not the contract
    want = orc.brute_hitcounts(r, hs)
    assert np.array_equal(c[ok], want[ok])
    assert np.array_equal(host(cum)[ok].view(np.uint32), orc.brute_cumulative(r, hs)[ok].view(np.uint32))


def test_hit_lists_with_sentinels(gb, orc, scene):
    d_s, tree, hs, htree, rays = scene
    rays = rays[:256]
    off = torch.empty(len(rays), dtype=torch.int32, device="cuda")
    idx, integ, dist = gb.trace_with_sentinels_sph(dev(rays), d_s, tree, off, -7, -1.0, 1e30)
    roff, ridx, rinteg, rdist = orc.trace_hits(rays, hs, htree)
    counts = np.diff(np.append(roff, len(ridx)))
    assert np.array_equal(host(off), roff + np.arange(len(rays)))
    assert len(idx) == len(ridx) + len(rays)
    hi = host(idx)
    for r in (0, 17, 255):
        b = roff[r] + r
        assert np.array_equal(hi[b:b + counts[r]], ridx[roff[r]:roff[r] + counts[r]])
        assert hi[b + counts[r]] == -7 and host(dist)[b + counts[r]] == np.float32(1e30)


def test_segsort_all_classes(gb, orc):
    # segment lengths straddling every size class incl. the global-memory path
    lens = [0, 1, 2, 31, 32, 33, 500, 512, 513, 1023, 1024, 1025, 1536, 1537, 2048, 2049, 3072, 3073,
            4096, 4097, 6144, 6145, 8192, 8193, 20000, 3, 0, 40000]
    rng = np.random.default_rng(3)
    total = sum(lens)
    dist = rng.integers(0, 1000, total).astype(np.float32) / np.float32(7.0)   # many ties
    dist[5] = 0.0
    dist[6] = -0.0
    idx = rng.integers(0, 1 << 30, total).astype(np.int32)
    data = rng.random(total, dtype=np.float32)
    off = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int32)
    d_dist, d_idx, d_data, d_off = dev(dist), dev(idx), dev(data), dev(off)
    gb.sort_by_distance(d_dist, d_off, d_idx, d_data)
    sd, si, sg = orc.sort_by_distance(dist, off, idx, data)
    assert np.array_equal(host(d_dist), sd)
    assert np.array_equal(host(d_idx), si)
    assert np.array_equal(host(d_data), sg)


def test_exclusive_scan(gb):
    rng = np.random.default_rng(0)
    for n in (1, 2047, 2048, 2049, 1000003):
        a = rng.integers(0, 2000, n).astype(np.int32)      # total < 2^31
        out, total = gb.exclusive_scan(dev(a))
        ref = np.concatenate([[0], np.cumsum(a[:-1], dtype=np.int64)])
        assert np.array_equal(host(out).astype(np.int64), ref)
        assert int(total.item()) == int(a.sum(dtype=np.int64))


def test_two_sphere_volume_integral(gb):
    # tests/integrate/integrate.cu:48-101 through the CUDA path
    radius = 0.2
    s = np.array([[-0.5, -0.5, -0.5, radius], [0.5, 0.5, 0.5, radius]], np.float32)
    d_s = dev(s)
    tree = gb.Tree(2, 1)
    gb.build_tree(d_s, tree, -np.ones(3, np.float32), np.ones(3, np.float32))
    n_side = 512
    span = 2.0 + 2 * radius
    rays = ortho_rays_z(n_side, -1.0 - radius, 1.0 + radius)
    rays[:, 5] = 1.0 + radius
    rays[:, 6] = 2 * span
    out = torch.empty(len(rays), dtype=torch.float32, device="cuda")
    gb.trace_cumulative_sph(dev(rays), d_s, tree, out)
    total = float(out.double().sum()) * (span / n_side) ** 2 / 2
    assert abs(1.0 - total) < 5e-4


def test_rays_not_multiple_of_32(gb, scene):
    d_s, tree, hs, htree, rays = scene
    out = torch.empty(33, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):           # bintree_trace.cuh:231-238
        gb.trace_hitcounts_sph(dev(rays[:33]), d_s, tree, out)


# ----------------------------------------------------------------------------- after the path
@pytest.mark.parametrize("count,max_seg,empty", [(200000, 3000, True), (50000, 20, False), (70000, 100000, True), (31, 5, True)])
def test_segmented_scans(gb, count, max_seg, empty):
    """tests/segmented_scan/segmented_scan.cu:66-160: random segments (empty ones allowed),
    integer-valued data -> exact equality with the sequential host scan; plus float data within
    rounding, the fused weights, offsets_to_segments (the reference's numbering) and in-place use."""
    rng = np.random.default_rng(count + max_seg)
    sizes = []
    total = 0
    while total < count:
        sz = int(rng.integers(0 if empty else 1, min(max_seg, count - total) + 1))
        sizes.append(sz)
        total += sz
    sizes = np.array(sizes, np.int64)
    off = (np.cumsum(sizes) - sizes).astype(np.int32)
    data = rng.integers(1, 10, count).astype(np.float32)
    weights = rng.integers(1, 5, 37).astype(np.float32)
    wmap = rng.integers(0, 37, count).astype(np.int32)
    seg_id = np.repeat(np.arange(len(sizes)), sizes)

    def host_scan(x):
        c = np.cumsum(x.astype(np.float64))
        start = np.concatenate([[0.0], c])[off.astype(np.int64)]
        return c - x - start[seg_id]
    res = torch.empty(count, dtype=torch.float32, device="cuda")
    gb.exclusive_segmented_scan(dev(off), dev(data), res)
    assert np.array_equal(host(res).astype(np.float64), host_scan(data))
    wres = torch.empty(count, dtype=torch.float32, device="cuda")
    gb.weighted_exclusive_segmented_scan(dev(data), dev(weights), dev(wmap), dev(off), wres)
    assert np.array_equal(host(wres).astype(np.float64), host_scan(weights[wmap] * data))
    inplace = dev(data)
    gb.exclusive_segmented_scan(dev(off), inplace, inplace)
    assert torch.equal(inplace, res)
    fdata = rng.random(count, dtype=np.float32)
    gb.exclusive_segmented_scan(dev(off), dev(fdata), res)
    want = host_scan(fdata)
    assert np.max(np.abs(host(res) - want) / np.maximum(want, 1.0)) < 1e-5
    segs = torch.zeros(count, dtype=torch.int32, device="cuda")
    gb.offsets_to_segments(dev(off), segs)
    flags = np.zeros(len(off), np.int64)
    if len(off) > 1:
        flags[1] = 1
        flags[2:] = off[2:] != off[1:-1]
    assert np.array_equal(host(segs), np.cumsum(flags)[seg_id])
    # calls sharing the context's device scalars must not disturb each other: bounds right after a segmented scan
    # (the scan's ticket counter used to be the bounds kernel's, which expects to find it zero)
    pts = uniform_spheres(70000, seed=9)
    lo, hi = gb.min_max_x(dev(pts))
    assert lo == float(pts[:, 0].min()) and hi == float(pts[:, 0].max())


@pytest.mark.parametrize("nside", [1, 4, 64, 1024])
def test_healpix_directions_bit_level(gb, orc, nside):
    """HEALPix NESTED pixel centres (RayVectorGeneration/src/chealpix/chealpix.c:112-126,357-391,459-467):
    the generator's directions against the oracle's restatement of pix2vec_nest rounded to float.  The double
    arithmetic is the same; sin / cos of the device library may differ from libm's in the last bit of a double,
    which survives rounding to float in well under one direction in a thousand -- those must be 1 ulp apart."""
    npix = 12 * nside * nside
    first = 0 if npix <= 4096 else 5 * nside * nside - 1000          # across a base-face boundary
    n = min(npix, 4096) // 32 * 32 or 32
    n = min(n, npix - first) // 32 * 32 if npix >= 32 else 0
    if n == 0:
        n, first = 0, 0
        # nside 1: 12 pixels -- rays must come in packets of 32 only for tracing, the generator takes any count
        rays = gb.healpix_rays(None, nside, 0, 12, 0.5, 0.25, 0.125, 3.0)
        n = 12
    else:
        rays = gb.healpix_rays(None, nside, first, n, 0.5, 0.25, 0.125, 3.0)
    got = host(rays)
    ref = orc.pix2vec_nest(nside, np.arange(first, first + n)).astype(np.float32)
    assert np.array_equal(got[:, 3:], np.tile(np.array([0.5, 0.25, 0.125, 3.0], np.float32), (n, 1)))
    d = np.abs(got[:, :3].view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    # +0.0 vs a value that rounds to the smallest denormals cannot happen here: |components| are 0 or >= 1e-7
    assert d.max() <= 1, d.max()
    assert (d != 0).mean() < 1e-3
    # unit length to float precision
    assert np.abs(np.linalg.norm(got[:, :3].astype(np.float64), axis=1) - 1.0).max() < 2e-7
