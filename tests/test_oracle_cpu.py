"""CPU-only checks that pin the oracle (oracle/) to the reference's own known-answer
tests and relational tests.  No GPU, no product code."""
import numpy as np
import pytest

from util import uniform_spheres, clustered_spheres, isotropic_rays, ortho_rays_z


def test_morton_30bit_kat(orc):
    # tests/morton_key/30bit_key.cu:20-26
    assert orc.space_by_two_10bit(309) == 16814145
    assert orc.space_by_two_10bit(942) == 153125448
    assert orc.space_by_two_10bit(619) == 134513161
    assert orc.morton_key30(309, 942, 619) == 861117685


def test_morton_63bit_kat(orc):
    # tests/morton_key/63bit_key.cu:20-26
    assert orc.morton_key63(1365301, 2014126, 1683051) == 8995068606879603957


def test_kernel_table_matches_analytic(orc):
    # LUT[i] = line integral of the Gadget-2 cubic spline at b/h = i/50.
    def W(q):
        q = np.asarray(q)
        return 8 / np.pi * np.where(q < 0.5, 1 - 6 * q ** 2 + 6 * q ** 3,
                                    np.where(q < 1, 2 * (1 - q) ** 3, 0.0))
    tab = orc.kernel_table()
    for i in (1, 10, 25, 49):
        b = i / 50.0
        zmax = np.sqrt(1 - b * b)
        z = np.linspace(0, zmax, 200001)
        f = W(np.sqrt(b * b + z * z))
        integral = 2 * np.sum((f[1:] + f[:-1]) * 0.5 * np.diff(z))
        assert abs(integral - tab[i]) < 2e-5 * max(1.0, tab[i])  # the table carries ~1e-5 quadrature error
    assert tab[50] == 0.0 and abs(tab[0] - 6 / np.pi) < 2e-6


def test_keys_are_device_formula(orc):
    s = uniform_spheres(5000, seed=3)
    bot = np.zeros(3, np.float32)
    top = np.ones(3, np.float32)
    k = orc.morton_keys(s, bot, top, 30)
    q = (np.float32(1023.0) / (top - bot) * (s[:, :3] - bot)).astype(np.uint32)
    ref = np.array([orc.morton_key30(*row) for row in q], np.uint32)
    assert np.array_equal(k, ref)
    k64 = orc.morton_keys(s, bot, top, 63)
    q = (np.float32(2097151.0) / (top - bot) * (s[:, :3] - bot)).astype(np.uint64)
    ref = np.array([orc.morton_key63(*row) for row in q], np.uint64)
    assert np.array_equal(k64, ref)


def test_sort_is_stable(orc):
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 50, 20000).astype(np.uint32)
    perm = orc.sort_perm(keys)
    assert np.array_equal(perm, np.argsort(keys, kind="stable"))
    keys64 = (rng.integers(0, 1 << 62, 20000).astype(np.uint64))
    assert np.array_equal(orc.sort_perm(keys64), np.argsort(keys64, kind="stable"))


def _leaves_by_definition(d, mpl):
    """Maximal subtrees with <= mpl primitives of the delta-defined Cartesian tree,
    ties to the right parent (albvh.cuh:129-143,181-186,269-291) -- independent scan form."""
    n = len(d) - 1
    dd = d[1:]
    out = []
    for j in range(n - 1):
        l = j
        while l > 0 and (j - l + 1) <= mpl and d[l] < dd[j]:
            l -= 1
        r = j + 1
        while r < n - 1 and (r - j) <= mpl and dd[r] <= dd[j]:
            r += 1
        ls, rs = j - l + 1, r - j
        big = ls + rs > mpl
        if big and ls <= mpl:
            out.append((l, ls))
        if big and rs <= mpl:
            out.append((j + 1, rs))
    return np.array(out, np.int32).reshape(-1, 2)


@pytest.mark.parametrize("mpl", [1, 2, 7, 32])
@pytest.mark.parametrize("kind", ["float", "ties", "u32", "u64"])
def test_leaf_clustering_two_formulations(orc, mpl, kind):
    rng = np.random.default_rng(mpl * 7 + len(kind))
    n = 2500
    if kind == "float":
        d = rng.random(n + 1).astype(np.float32)
        d[0] = d[-1] = np.inf
    elif kind == "ties":
        d = rng.integers(0, 4, n + 1).astype(np.float32)
        d[0] = d[-1] = np.inf
    elif kind == "u32":
        d = rng.integers(0, 60, n + 1).astype(np.uint32)
        d[0] = d[-1] = 0xFFFFFFFF
    else:
        d = rng.integers(0, 1 << 40, n + 1).astype(np.uint64)
        d[0] = d[-1] = 0xFFFFFFFFFFFFFFFF
    leaves = orc.build_leaves(d, mpl)
    assert np.array_equal(leaves[:, :2], _leaves_by_definition(d, mpl))
    # leaves tile [0, n) and respect the cap
    assert leaves[0, 0] == 0 and leaves[-1, 0] + leaves[-1, 1] == n
    assert np.array_equal(leaves[1:, 0], leaves[:-1, 0] + leaves[:-1, 1])
    assert leaves[:, 1].max() <= mpl


def _check_tree(tree, spheres):
    L = tree.n_leaves
    nn = L - 1
    nodes = tree.nodes
    fn = nodes.view(np.float32)
    seen_child = np.zeros(nn + L, bool)
    for j in range(nn):
        left, right, first, last = nodes[j, :4]
        for c in (left, right):
            assert not seen_child[c]
            seen_child[c] = True
        # node index = split position: left covers [first, j], right covers [j+1, last]
        assert first <= j < last
    assert seen_child.sum() == nn + L - 1 and not seen_child[tree.root]
    assert nodes[tree.root, 2] == 0 and nodes[tree.root, 3] == L - 1
    # boxes contain their primitives
    lo = spheres[:, :3] - spheres[:, 3:4]
    hi = spheres[:, :3] + spheres[:, 3:4]
    for j in range(0, nn, max(1, nn // 200)):
        for side, child, cols in ((0, nodes[j, 0], (4, 5, 6, 7, 12, 13)), (1, nodes[j, 1], (8, 9, 10, 11, 14, 15))):
            if child >= nn:
                a = b = child - nn
            else:
                a, b = nodes[child, 2], nodes[child, 3]
            p0 = tree.leaves[a, 0]
            p1 = tree.leaves[b, 0] + tree.leaves[b, 1]
            box = fn[j, list(cols)]
            assert np.all(box[[0, 2, 4]] == lo[p0:p1].min(0))
            assert np.all(box[[1, 3, 5]] == hi[p0:p1].max(0))


@pytest.mark.parametrize("mpl", [1, 8, 32])
def test_tree_structure(orc, mpl):
    s = clustered_spheres(20000, seed=5)
    ss, keys, perm = orc.sort_spheres(s, 30)
    tree = orc.build_tree(ss, orc.deltas_euclid(ss), mpl)
    _check_tree(tree, ss)
    with pytest.raises(ValueError):
        orc.build_tree(ss[:mpl], orc.deltas_euclid(ss[:mpl]), mpl)


def test_tree_xor_deltas(orc):
    s = uniform_spheres(30000, seed=9)
    ss, keys, perm = orc.sort_spheres(s, 63)
    tree = orc.build_tree(ss, orc.deltas_xor(keys), 16)
    _check_tree(tree, ss)


def test_trace_equals_brute_force(orc):
    # The reference's own oracle pattern: tests/tree_traversal/tree_traversal.cu:65-121.
    s = uniform_spheres(1 << 14, seed=11, rmax=0.05)
    ss, _, _ = orc.sort_spheres(s, 30, np.zeros(3, np.float32), np.ones(3, np.float32))
    tree = orc.build_tree(ss, orc.deltas_euclid(ss), 32)
    rays = isotropic_rays(1 << 11, seed=2)
    counts = orc.trace_hitcounts(rays, ss, tree)
    assert np.array_equal(counts, orc.brute_hitcounts(rays, ss))
    # cumulative sums run in ascending primitive order in both (left-first DFS)
    assert np.array_equal(orc.trace_cumulative(rays, ss, tree), orc.brute_cumulative(rays, ss))


def test_two_sphere_volume_integral(orc):
    # tests/integrate/integrate.cu:48-101: sum(column density) * pixel area / N == 1 +- 5e-4
    radius = 0.2
    s = np.array([[-0.5, -0.5, -0.5, radius], [0.5, 0.5, 0.5, radius]], np.float32)
    ss, _, _ = orc.sort_spheres(s, 30, -np.ones(3, np.float32), np.ones(3, np.float32))
    tree = orc.build_tree(ss, orc.deltas_euclid(ss), 1)
    n_side = 512
    span = 2.0 + 2 * radius
    rays = ortho_rays_z(n_side, -1.0 - radius, 1.0 + radius)
    rays[:, 5] = 1.0 + radius
    rays[:, 6] = 2 * span
    cum = orc.trace_cumulative(rays, ss, tree)
    total = cum.astype(np.float64).sum() * (span / n_side) ** 2 / 2
    assert abs(1.0 - total) < 5e-4


def test_hit_lists_sorted(orc):
    # tests/distance_sort/distance_sort.cu:22-79
    s = uniform_spheres(20000, seed=4, rmax=0.05)
    ss, _, _ = orc.sort_spheres(s, 30)
    tree = orc.build_tree(ss, orc.deltas_euclid(ss), 32)
    rays = isotropic_rays(512, seed=6)
    off, idx, integ, dist = orc.trace_hits(rays, ss, tree)
    d2, i2, g2 = orc.sort_by_distance(dist, off, idx, integ)
    ends = np.append(off[1:], len(dist))
    for b, e in zip(off, ends):
        seg = d2[b:e]
        assert np.all(seg >= 0) and np.all(np.diff(seg) >= 0)
        assert sorted(i2[b:e]) == sorted(idx[b:e])
    # stable: equals numpy's stable argsort per segment
    b, e = off[3], ends[3]
    order = np.argsort(dist[b:e], kind="stable")
    assert np.array_equal(i2[b:e], idx[b:e][order])


def test_healpix_unit_vectors(orc):
    nside = 8
    v = orc.pix2vec_nest(nside, np.arange(12 * nside * nside))
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-14)
    # equal-area pixelisation: mean vector ~ 0, and all pixel centres distinct
    assert np.abs(v.mean(0)).max() < 1e-12
    assert len(np.unique(np.round(v, 12), axis=0)) == len(v)
    # pixel 0 of face 0 at nside=1 is (theta = acos(2/3), phi = pi/4)
    v1 = orc.pix2vec_nest(1, [0])[0]
    assert abs(v1[2] - 2 / 3) < 1e-15 and abs(v1[0] - v1[1]) < 1e-15
