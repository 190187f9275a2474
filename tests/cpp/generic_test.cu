// generic_test.cu -- GRACE's generic (user-defined primitive + functor) path against this repo's
// header templates, shaped like the reference's triangle demo
// (tests/profile_trace_triangle/tris_tree.cuh:17-30, tris_trace.cu:46-66): Morton keys from a
// user centroid functor, sort, XOR deltas, build_ALBVH with a user AABB functor, closest-hit
// trace with user functors.  Checks against host brute force; with an output directory it also
// dumps the arrays tests/test_gpu_generic.py compares with the reference's own build.
//   generic_test [n_tris=200000] [n_rays=65536] [outdir]
#include "grace/cuda/functors/trace.cuh"
#include "grace/cuda/kernels/albvh.cuh"
#include "grace/cuda/kernels/bintree_trace.cuh"
#include "grace/cuda/kernels/morton.cuh"
#include "grace/cuda/sort_by_key.cuh"
#include "grace/generic/functors/albvh.h"

#include "generic_prims.cuh"

#include <cstdio>
#include <cstdlib>
#include <string>

template <typename T>
static void dump(const std::string& dir, const char* name, const std::vector<T>& v)
{
    if (dir.empty()) return;
    FILE* f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror(name); exit(2); }
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

int main(int argc, char** argv)
{
    const size_t N = argc > 1 ? strtoul(argv[1], 0, 10) : 200000;
    const size_t R = argc > 2 ? strtoul(argv[2], 0, 10) : 65536;
    const std::string out = argc > 3 ? argv[3] : "";
    const int max_per_leaf = 8;
    const std::vector<Tri> h_tris = make_tris(N, 0.02f);
    const std::vector<grace::Ray> h_rays = make_rays(R);

    grace::device_vector<Tri> d_tris(h_tris);
    grace::device_vector<grace::uinteger32> d_keys(N), d_deltas(N + 1);
    float3 bot, top;
    grace::morton_keys(d_tris, d_keys, TriCentroid(), &bot, &top);
    grace::sort_by_key(d_keys, d_tris);
    grace::compute_deltas(d_keys, d_deltas, grace::DeltaXOR());
    grace::Tree d_tree(N, max_per_leaf);
    grace::build_ALBVH(d_tree, d_tris, d_deltas, TriAABB());

    grace::device_vector<grace::Ray> d_rays(h_rays);
    grace::device_vector<int> d_closest(R), d_counts(R);
    grace::trace_texref<RayData_tri>(d_rays, d_tris, d_tree, 0, grace::Init_null(), RayIntersect_tri(), OnHit_tri(),
                                     RayEntry_tri(), grace::RayExit_to_array<int>(d_closest.data()));
    grace::trace<RayData_cnt>(d_rays, d_tris, d_tree, 0, grace::Init_null(), RayIntersect_any(), grace::OnHit_increment(),
                              grace::RayEntry_null(), grace::RayExit_to_array<int>(d_counts.data()));
    const std::vector<Tri> s_tris = d_tris.to_host();
    const std::vector<int> closest = d_closest.to_host(), counts = d_counts.to_host();

    // host brute force on a sample of rays: counts exact; closest: same t (ties may pick another index)
    size_t bad = 0, checked = 0;
    for (size_t r = 0; r < R; r += R / 512 ? R / 512 : 1, ++checked) {
        int cnt = 0, best = -1;
        float tmin = h_rays[r].length;
        for (size_t i = 0; i < N; ++i) {
            float t;
            if (!tri_intersect(h_rays[r], s_tris[i], &t)) continue;
            if (t >= 0.f && t < h_rays[r].length) ++cnt;
            if (t <= tmin && t >= 1e-6f) { if (t < tmin || best < 0) best = (int)i; tmin = t; }
        }
        if (cnt != counts[r]) ++bad;
        float tg = -1.f;
        if (closest[r] >= 0) tri_intersect(h_rays[r], s_tris[closest[r]], &tg);
        if ((best < 0) != (closest[r] < 0) || (best >= 0 && tg != tmin)) ++bad;
    }
    long long total = 0;
    for (int c : counts) total += c;
    std::printf("generic: %zu triangles, %zu leaves, %zu rays, %lld crossings, %zu sampled rays, %zu mismatches vs host brute force\n",
                N, d_tree.leaves.size(), R, total, checked, bad);
    dump(out, "keys.bin", d_keys.to_host());
    dump(out, "tris.bin", s_tris);
    dump(out, "deltas.bin", d_deltas.to_host());
    dump(out, "leaves.bin", d_tree.leaves.to_host());
    dump(out, "nodes.bin", d_tree.nodes.to_host());
    dump(out, "closest.bin", closest);
    dump(out, "counts.bin", counts);
    int root = 0;
    cudaMemcpy(&root, d_tree.root_index_ptr, sizeof(int), cudaMemcpyDeviceToHost);
    dump(out, "root.bin", std::vector<int>(1, root));
    std::printf("%s generic primitives\n", bad == 0 && total > 0 ? "PASSED" : "FAILED");
    return bad == 0 && total > 0 ? 0 : 1;
}
