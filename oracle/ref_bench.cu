// ref_bench.cu -- the bench workload (bench.py, BASELINE.json configs[2] "profile_trace_gadget")
// run through the REFERENCE's own public API (GRACE, patched only for CUDA-12 compatibility by
// oracle/patch_ref.py): the reference arm of bench.py and its "reference CUDA build on the same
// box" comparator.  Built twice by oracle/build_ref.sh: as shipped
// (kernel_config.h:11 MAX_BLOCKS = 112) and "tuned" (MAX_BLOCKS = 148 x 8).
//
// TEST / MEASUREMENT INFRASTRUCTURE: compiled from the sources under /root/reference (never copied
// into this repository) into oracle/_ref/.  The only code shared with the product is the synthetic
// snapshot generator (grace-devel_b200/csrc/synth.cu, linked as an object file -- libgrace_b200.so
// is not loaded), so that both arms trace the same particles; everything timed is the reference's.
//
//   ref_bench <log2 particles> <log2 rays> <max_per_leaf> <steps> <warmup> [e2e steps [dump dir [sample rays]]]
// With a dump dir: spheres_sorted.bin (the particles as the trace sees them), rays_sample.bin and
// cum_sample.bin (`sample rays` evenly strided rays and their column densities) for the host
// brute-force comparison.
// One step = grace::trace_cumulative_sph over all rays (device-resident inputs, CUDA events, a
// 256 MiB write between steps to flush L2).  One e2e step = particles H2D, tree build
// (morton_keys30_sort_sph, euclidean_deltas_sph, ALBVH_sph: tests/helper/tree.cuh:15-27), rays H2D,
// trace, result D2H.  Prints one JSON object.
#include <curand_kernel.h>

#include "grace/cuda/nodes.h"
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/gen_rays.cuh"
#include "grace/cuda/trace_sph.cuh"
#include "grace/cuda/util/extrema.cuh"
#include "grace/ray.h"

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "grace_b200.h"     // grace_b200_create / _synth_gadget_f4 only (workload generator)

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

static unsigned long long fnv1a(const void* p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}

static void build(thrust::device_vector<float4>& d_spheres, thrust::device_vector<float>& d_deltas, grace::Tree& tree)
{
    grace::morton_keys30_sort_sph(d_spheres);
    grace::euclidean_deltas_sph(d_spheres, d_deltas);
    grace::ALBVH_sph(d_spheres, d_deltas, tree);
}

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: ref_bench log2N log2R max_per_leaf steps warmup [e2e_steps]\n"); return 2; }
    const size_t N = (size_t)1 << atoi(argv[1]), R = (size_t)1 << atoi(argv[2]);
    const int mpl = atoi(argv[3]), steps = atoi(argv[4]), warmup = atoi(argv[5]);
    const int e2e_steps = argc > 6 ? atoi(argv[6]) : steps;

    // ---- the same synthetic snapshot as the other arm ----
    float4* h_in = nullptr;
    cudaMallocHost((void**)&h_in, N * sizeof(float4));
    {
        grace_b200_ctx* ctx = nullptr;
        if (grace_b200_create(&ctx, 0)) { fprintf(stderr, "%s\n", grace_b200_last_error()); return 3; }
        float* d = nullptr;
        cudaMalloc((void**)&d, N * sizeof(float4));
        if (grace_b200_synth_gadget_f4(ctx, d, N, 1234u, nullptr)) { fprintf(stderr, "%s\n", grace_b200_last_error()); return 3; }
        cudaMemcpy(h_in, d, N * sizeof(float4), cudaMemcpyDeviceToHost);
        cudaFree(d);
        grace_b200_destroy(ctx);
    }

    Timer T;
    thrust::device_vector<float4> d_spheres(h_in, h_in + N);
    thrust::device_vector<float> d_deltas(N + 1);
    grace::Tree* tree = new grace::Tree(N, mpl);
    build(d_spheres, d_deltas, *tree);
    const size_t n_leaves = tree->leaves.size();

    float lo, hi;
    grace::min_max_x(d_spheres, &lo, &hi);
    const float c = (hi + lo) / 2.0f, len = 2.0f * (hi - lo);       // profile_trace_gadget.cu:82-84
    thrust::device_vector<grace::Ray> d_rays(R);
    grace::uniform_random_rays(d_rays, c, c, c, len, 1234ull);
    grace::Ray* h_rays = nullptr;
    cudaMallocHost((void**)&h_rays, R * sizeof(grace::Ray));
    cudaMemcpy(h_rays, thrust::raw_pointer_cast(d_rays.data()), R * sizeof(grace::Ray), cudaMemcpyDeviceToHost);

    thrust::device_vector<float> d_cum(R);
    float* h_cum = nullptr;
    cudaMallocHost((void**)&h_cum, R * sizeof(float));
    void* flush = nullptr;
    cudaMalloc(&flush, 256u << 20);

    for (int i = 0; i < warmup; ++i) grace::trace_cumulative_sph(d_rays, d_spheres, *tree, d_cum);
    cudaDeviceSynchronize();
    double sum = 0, best = 1e30;
    for (int i = 0; i < steps; ++i) {
        cudaMemsetAsync(flush, i & 0xff, 256u << 20);
        T.start();
        grace::trace_cumulative_sph(d_rays, d_spheres, *tree, d_cum);
        const float ms = T.stop();
        sum += ms; if (ms < best) best = ms;
    }
    cudaMemcpy(h_cum, thrust::raw_pointer_cast(d_cum.data()), R * sizeof(float), cudaMemcpyDeviceToHost);
    const unsigned long long sha = fnv1a(h_cum, R * sizeof(float));

    // ---- end to end, from host buffers ----
    double e_sum = 0;
    for (int i = -1; i < e2e_steps; ++i) {
        cudaMemsetAsync(flush, i & 0xff, 256u << 20);
        cudaDeviceSynchronize();
        T.start();
        cudaMemcpyAsync(thrust::raw_pointer_cast(d_spheres.data()), h_in, N * sizeof(float4), cudaMemcpyHostToDevice);
        delete tree;
        tree = new grace::Tree(N, mpl);
        build(d_spheres, d_deltas, *tree);
        cudaMemcpyAsync(thrust::raw_pointer_cast(d_rays.data()), h_rays, R * sizeof(grace::Ray), cudaMemcpyHostToDevice);
        grace::trace_cumulative_sph(d_rays, d_spheres, *tree, d_cum);
        cudaMemcpyAsync(h_cum, thrust::raw_pointer_cast(d_cum.data()), R * sizeof(float), cudaMemcpyDeviceToHost);
        const float ms = T.stop();
        if (i >= 0) e_sum += ms;
    }
    const unsigned long long sha2 = fnv1a(h_cum, R * sizeof(float));
    if (argc > 7) {
        const std::string dir = argv[7];
        const size_t S = argc > 8 ? (size_t)atol(argv[8]) : 1024;
        std::vector<float4> hs(N);
        cudaMemcpy(hs.data(), thrust::raw_pointer_cast(d_spheres.data()), N * sizeof(float4), cudaMemcpyDeviceToHost);
        FILE* f = fopen((dir + "/spheres_sorted.bin").c_str(), "wb");
        if (f) { fwrite(hs.data(), sizeof(float4), N, f); fclose(f); }
        std::vector<grace::Ray> rs(S);
        std::vector<float> cs(S);
        const size_t stride = R / S ? R / S : 1;
        for (size_t i = 0; i < S; ++i) { rs[i] = h_rays[(i * stride) % R]; cs[i] = h_cum[(i * stride) % R]; }
        f = fopen((dir + "/rays_sample.bin").c_str(), "wb");
        if (f) { fwrite(rs.data(), sizeof(grace::Ray), S, f); fclose(f); }
        f = fopen((dir + "/cum_sample.bin").c_str(), "wb");
        if (f) { fwrite(cs.data(), sizeof(float), S, f); fclose(f); }
    }
    printf("{\"impl\": \"reference-cuda\", \"max_blocks\": %d, \"particles\": %zu, \"rays\": %zu, \"n_leaves\": %zu, "
           "\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.4f, \"best_ms\": %.4f, \"e2e_steps\": %d, \"e2e_ms_per_step\": %.4f, "
           "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, \"result_fnv1a\": \"%016llx\", \"e2e_result_fnv1a\": \"%016llx\"}\n",
           (int)grace::MAX_BLOCKS, N, R, n_leaves, steps, warmup, steps ? sum / steps : 0.0, best, e2e_steps,
           e2e_steps ? e_sum / e2e_steps : 0.0, N * sizeof(float4) + R * sizeof(grace::Ray), R * sizeof(float), sha, sha2);
    delete tree;
    return 0;
}
