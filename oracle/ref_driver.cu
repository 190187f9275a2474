// ref_driver.cu -- runs the REFERENCE (GRACE, patched only for CUDA-12 compatibility by
// oracle/patch_ref.py) through its own public API on given inputs, dumps every intermediate
// of the hot path as raw binary and prints per-stage timings.
//
// TEST INFRASTRUCTURE: compiled by oracle/build_ref.sh into oracle/_ref/ref_driver from the
// sources under /root/reference (never copied into this repository).  Used (i) to pin the
// CPU oracle and the CUDA product bit-for-bit to the reference's own CUDA implementation
// (tests/test_gpu_vs_reference.py, tests/golden/), (ii) as the "reference CUDA build on the
// same box" comparator of BASELINE.md 2b.
//
//   ref_driver <spheres.bin> <rays.bin | gen:N:seed:ox:oy:oz:len> <outdir> <max_per_leaf>
//              <key_bits 30|63> <iters> [lists]
// spheres.bin: N x float4; rays.bin: R x 7 floats.  Outputs in <outdir>:
//   spheres_sorted.bin, deltas.bin, leaves.bin (L x int4), nodes.bin (4(L-1) x int4),
//   root.bin, rays.bin, hitcounts.bin, cumulative.bin and, with "lists": offsets.bin,
//   hit_idx.bin, hit_integral.bin, hit_dist.bin (sorted by distance).
#include <curand_kernel.h>

#include "grace/cuda/nodes.h"
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/gen_rays.cuh"
#include "grace/cuda/trace_sph.cuh"
#include "grace/cuda/sort.cuh"
#include "grace/cuda/util/extrema.cuh"
#include "grace/ray.h"

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

template <typename T>
static void dump(const std::string& dir, const char* name, const T* p, size_t n)
{
    FILE* f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror(name); exit(2); }
    fwrite(p, sizeof(T), n, f);
    fclose(f);
}
template <typename T>
static void dump(const std::string& dir, const char* name, const thrust::device_vector<T>& d)
{
    thrust::host_vector<T> h = d;
    dump(dir, name, thrust::raw_pointer_cast(h.data()), h.size());
}
template <typename T>
static std::vector<T> slurp(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v(bytes / sizeof(T));
    if (fread(v.data(), sizeof(T), v.size(), f) != v.size()) { perror("read"); exit(2); }
    fclose(f);
    return v;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: see header\n"); return 2; }
    const std::string out = argv[3];
    const int max_per_leaf = atoi(argv[4]);
    const int key_bits = atoi(argv[5]);
    const int iters = atoi(argv[6]);
    const bool lists = argc > 7 && !strcmp(argv[7], "lists");

    std::vector<float4> h_in = slurp<float4>(argv[1]);
    const size_t N = h_in.size();
    Timer T;
    double t_sort = 0, t_deltas = 0, t_build = 0, t_hit = 0, t_cum = 0, t_lists = 0, t_sortd = 0, t_gen = 0;
    // best iteration per stage: the reference allocates its temporaries inside every call, which
    // makes single iterations noisy (a GiB-sized cudaMalloc can take longer than the kernels)
    double m_sort = 1e30, m_deltas = 1e30, m_build = 1e30, m_hit = 1e30, m_cum = 1e30, m_lists = 1e30, m_sortd = 1e30, m_gen = 1e30;
#define ACC(sum, mn) if (it >= 0) { sum += ms; if (ms < mn) mn = ms; }
    size_t n_leaves = 0;

    thrust::device_vector<float4> d_spheres;
    thrust::device_vector<float> d_deltas(N + 1);
    grace::Tree* tree = NULL;
    for (int it = -1; it < iters; ++it) {          // it == -1: warm-up (profile_tree_gadget.cu:85)
        d_spheres = h_in;
        delete tree;
        tree = new grace::Tree(N, max_per_leaf);
        cudaDeviceSynchronize();
        T.start();
        if (key_bits == 30) grace::morton_keys30_sort_sph(d_spheres);
        else grace::morton_keys63_sort_sph(d_spheres);
        float ms = T.stop(); ACC(t_sort, m_sort)
        T.start();
        grace::euclidean_deltas_sph(d_spheres, d_deltas);
        ms = T.stop(); ACC(t_deltas, m_deltas)
        T.start();
        grace::ALBVH_sph(d_spheres, d_deltas, *tree);
        ms = T.stop(); ACC(t_build, m_build)
        n_leaves = tree->leaves.size();
    }
    dump(out, "spheres_sorted.bin", d_spheres);
    dump(out, "deltas.bin", d_deltas);
    dump(out, "leaves.bin", tree->leaves);
    {
        thrust::host_vector<int4> h_nodes(tree->nodes.begin(), tree->nodes.begin() + 4 * (n_leaves - 1));
        dump(out, "nodes.bin", thrust::raw_pointer_cast(h_nodes.data()), h_nodes.size());
    }
    int root;
    cudaMemcpy(&root, tree->root_index_ptr, sizeof(int), cudaMemcpyDeviceToHost);
    dump(out, "root.bin", &root, 1);

    // ---- rays ----
    thrust::device_vector<grace::Ray> d_rays;
    if (!strncmp(argv[2], "gen:", 4)) {
        size_t R; unsigned long long seed; float ox, oy, oz, len;
        if (sscanf(argv[2] + 4, "%zu:%llu:%f:%f:%f:%f", &R, &seed, &ox, &oy, &oz, &len) != 6) return 2;
        d_rays.resize(R);
        for (int it = -1; it < iters; ++it) {
            T.start();
            grace::uniform_random_rays(d_rays, ox, oy, oz, len, seed);
            float ms = T.stop(); ACC(t_gen, m_gen)
        }
    } else {
        std::vector<grace::Ray> h_rays = slurp<grace::Ray>(argv[2]);
        d_rays = thrust::host_vector<grace::Ray>(h_rays.begin(), h_rays.end());
    }
    const size_t R = d_rays.size();
    dump(out, "rays.bin", d_rays);

    thrust::device_vector<int> d_counts(R);
    thrust::device_vector<float> d_cum(R);
    for (int it = -1; it < iters; ++it) {
        T.start();
        grace::trace_hitcounts_sph(d_rays, d_spheres, *tree, d_counts);
        float ms = T.stop(); ACC(t_hit, m_hit)
        T.start();
        grace::trace_cumulative_sph(d_rays, d_spheres, *tree, d_cum);
        ms = T.stop(); ACC(t_cum, m_cum)
    }
    dump(out, "hitcounts.bin", d_counts);
    dump(out, "cumulative.bin", d_cum);

    size_t total_hits = 0;
    if (lists) {
        thrust::device_vector<int> d_offsets(R), d_idx;
        thrust::device_vector<float> d_integ, d_dist;
        for (int it = -1; it < iters; ++it) {
            T.start();
            grace::trace_sph(d_rays, d_spheres, *tree, d_offsets, d_idx, d_integ, d_dist);
            float ms = T.stop(); ACC(t_lists, m_lists)
            T.start();
            grace::sort_by_distance(d_dist, d_offsets, d_idx, d_integ);
            ms = T.stop(); ACC(t_sortd, m_sortd)
        }
        total_hits = d_idx.size();
        dump(out, "offsets.bin", d_offsets);
        dump(out, "hit_idx.bin", d_idx);
        dump(out, "hit_integral.bin", d_integ);
        dump(out, "hit_dist.bin", d_dist);
    }
    const double k = iters > 0 ? 1.0 / iters : 0.0;
    printf("{\"impl\": \"reference-cuda\", \"n\": %zu, \"rays\": %zu, \"n_leaves\": %zu, \"root\": %d, "
           "\"max_per_leaf\": %d, \"key_bits\": %d, \"iters\": %d, "
           "\"ms_keys_sort\": %.4f, \"ms_deltas\": %.4f, \"ms_albvh\": %.4f, \"ms_gen_rays\": %.4f, "
           "\"ms_hitcounts\": %.4f, \"ms_cumulative\": %.4f, \"ms_trace_lists\": %.4f, "
           "\"ms_sort_by_distance\": %.4f, \"total_hits\": %zu, "
           "\"min_ms\": {\"keys_sort\": %.4f, \"deltas\": %.4f, \"albvh\": %.4f, \"gen_rays\": %.4f, \"hitcounts\": %.4f, "
           "\"cumulative\": %.4f, \"trace_lists\": %.4f, \"sort_by_distance\": %.4f}}\n",
           N, R, n_leaves, root, max_per_leaf, key_bits, iters, t_sort * k, t_deltas * k, t_build * k,
           t_gen * k, t_hit * k, t_cum * k, t_lists * k, t_sortd * k, total_hits,
           m_sort, m_deltas, m_build, m_gen < 1e29 ? m_gen : 0.0, m_hit, m_cum, m_lists < 1e29 ? m_lists : 0.0,
           m_sortd < 1e29 ? m_sortd : 0.0);
    delete tree;
    return 0;
}
