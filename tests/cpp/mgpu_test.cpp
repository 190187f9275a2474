// mgpu_test.cpp -- the multi-GPU C ABI (include/grace_b200_mgpu.h) against the single-GPU one:
// a tree built on every device / built on device 0 and broadcast is the same tree, bit for bit, and
// rays dealt over the devices in 32-aligned tiles give the outputs of a one-GPU trace, bit for bit
// (SURVEY.md 8e).  Runs on however many devices are visible (1 included); plain C++, no Python.
//
//   mgpu_test [log2 particles = 20] [rays = 3 tiles + 1 packet short of 16 tiles] [devices = all]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "grace_b200_mgpu.h"

#define CHECK(call) do { int rc_ = (call); if (rc_) { std::printf("FAILED %s -> %d: %s | %s\n", #call, rc_, grace_b200_last_error(), grace_b200_mgpu_last_error()); return 1; } } while (0)

int main(int argc, char** argv)
{
    const size_t n = (size_t)1 << (argc > 1 ? std::atoi(argv[1]) : 20);
    const size_t n_rays = argc > 2 ? (size_t)std::atol(argv[2]) : 16 * 4096 - 4096 + 2048 + 32;   // ragged last tile
    const int want = argc > 3 ? std::atoi(argv[3]) : 0;
    const int mpl = 32;

    // ---- inputs and the single-GPU answers (device 0, plain grace_b200) ----
    grace_b200_ctx* ctx = nullptr;
    CHECK(grace_b200_create(&ctx, 0));
    float *d_s = nullptr, *d_raw = nullptr;
    cudaMalloc((void**)&d_raw, n * 16);
    cudaMalloc((void**)&d_s, n * 16);
    CHECK(grace_b200_synth_gadget_f4(ctx, d_raw, n, 1234u, nullptr));
    std::vector<float> h_raw(4 * n);
    cudaMemcpy(h_raw.data(), d_raw, n * 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(d_s, d_raw, n * 16, cudaMemcpyDeviceToDevice);
    CHECK(grace_b200_morton_sort_f4(ctx, d_s, n, 30, nullptr, nullptr, nullptr, nullptr));
    float* d_deltas = nullptr;
    cudaMalloc((void**)&d_deltas, (n + 1) * 4);
    CHECK(grace_b200_deltas_euclid_f4(ctx, d_s, n, d_deltas, nullptr));
    void *d_nodes = nullptr, *d_leaves = nullptr;
    int* d_root = nullptr;
    cudaMalloc(&d_nodes, 64 * (n - 1));
    cudaMalloc(&d_leaves, 16 * n);
    cudaMalloc((void**)&d_root, 4);
    int L = 0;
    CHECK(grace_b200_albvh_build_f4(ctx, d_s, n, d_deltas, GRACE_B200_DELTA_F32, mpl, d_nodes, d_leaves, d_root, &L, nullptr));
    float mm[8];
    CHECK(grace_b200_minmax_f4_host(ctx, d_s, n, mm, nullptr));
    const float c = 0.5f * (mm[0] + mm[4]), len = 2.f * (mm[4] - mm[0]);
    grace_b200_ray* d_rays = nullptr;
    cudaMalloc((void**)&d_rays, n_rays * sizeof(grace_b200_ray));
    CHECK(grace_b200_uniform_random_rays(ctx, d_rays, n_rays, c, c, c, len, -1, 1234ull, nullptr));
    std::vector<grace_b200_ray> h_rays(n_rays);
    cudaMemcpy(h_rays.data(), d_rays, n_rays * sizeof(grace_b200_ray), cudaMemcpyDeviceToHost);
    grace_b200_tree tr = { d_nodes, d_leaves, d_root, L, mpl };
    float* d_cum = nullptr; int* d_cnt = nullptr;
    cudaMalloc((void**)&d_cum, n_rays * 4);
    cudaMalloc((void**)&d_cnt, n_rays * 4);
    CHECK(grace_b200_trace_cumulative_f4(ctx, d_rays, n_rays, d_s, n, &tr, d_cum, nullptr));
    CHECK(grace_b200_trace_hitcounts_f4(ctx, d_rays, n_rays, d_s, n, &tr, d_cnt, nullptr));
    std::vector<float> ref_cum(n_rays); std::vector<int> ref_cnt(n_rays);
    cudaMemcpy(ref_cum.data(), d_cum, n_rays * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref_cnt.data(), d_cnt, n_rays * 4, cudaMemcpyDeviceToHost);
    std::vector<int> ref_nodes(16 * (size_t)(L - 1));
    std::vector<float> ref_s(4 * n);
    cudaMemcpy(ref_nodes.data(), d_nodes, 64 * (size_t)(L - 1), cudaMemcpyDeviceToHost);
    cudaMemcpy(ref_s.data(), d_s, 16 * n, cudaMemcpyDeviceToHost);

    // ---- the multi-GPU layer ----
    grace_b200_mgpu* mg = nullptr;
    CHECK(grace_b200_mgpu_init(&mg, want, nullptr));
    const int world = grace_b200_mgpu_n_devices(mg);
    int failures = 0;
    for (int how = 0; how < 2; ++how) {
        int L2 = 0;
        float ms3[3], ms4[4];
        CHECK(grace_b200_mgpu_build_f4(mg, h_raw.data(), n, mpl, 30, how, &L2, ms3));
        bool same_tree = L2 == L;
        for (int dev = 0; dev < world && same_tree; dev += (world > 1 ? world - 1 : 1)) {       // first and last device
            std::vector<int> nodes(16 * (size_t)(L - 1)), leaves(4 * (size_t)L);
            std::vector<float> s(4 * n);
            int root = -1, ref_root = -2;
            CHECK(grace_b200_mgpu_copy_tree(mg, dev, s.data(), nodes.data(), leaves.data(), &root));
            cudaMemcpy(&ref_root, d_root, 4, cudaMemcpyDeviceToHost);
            same_tree = root == ref_root && !std::memcmp(nodes.data(), ref_nodes.data(), nodes.size() * 4) &&
                        !std::memcmp(s.data(), ref_s.data(), s.size() * 4);
        }
        std::vector<float> cum(n_rays); std::vector<int> cnt(n_rays);
        CHECK(grace_b200_mgpu_trace_cumulative_f4(mg, h_rays.data(), n_rays, cum.data(), ms4));
        CHECK(grace_b200_mgpu_trace_hitcounts_f4(mg, h_rays.data(), n_rays, cnt.data(), nullptr));
        const bool same_cum = !std::memcmp(cum.data(), ref_cum.data(), n_rays * 4);
        const bool same_cnt = !std::memcmp(cnt.data(), ref_cnt.data(), n_rays * 4);
        std::printf("{\"devices\": %d, \"build\": \"%s\", \"n\": %zu, \"rays\": %zu, \"n_leaves\": %d, \"tree_bit_identical\": %s, "
                    "\"column_densities_bit_identical\": %s, \"hit_counts_identical\": %s, \"ms_h2d\": %.3f, \"ms_broadcast\": %.3f, "
                    "\"ms_build\": %.3f, \"ms_rays_in\": %.3f, \"ms_trace\": %.3f, \"ms_gather\": %.3f, \"ms_result_out\": %.3f}\n",
                    world, how ? "on device 0, tree broadcast" : "on every device", n, n_rays, L2, same_tree ? "true" : "false",
                    same_cum ? "true" : "false", same_cnt ? "true" : "false", ms3[0], ms3[1], ms3[2], ms4[0], ms4[1], ms4[2], ms4[3]);
        failures += !same_tree + !same_cum + !same_cnt;
    }
    // argument errors as in the single-GPU API (bintree_trace.cuh:231-238)
    {
        float dummy[64];
        if (grace_b200_mgpu_trace_cumulative_f4(mg, h_rays.data(), 33, dummy, nullptr) != GRACE_B200_EINVAL) { std::printf("FAILED: 33 rays accepted\n"); ++failures; }
    }
    CHECK(grace_b200_mgpu_finalize(mg));
    grace_b200_destroy(ctx);
    std::printf(failures ? "FAILED\n" : "PASSED\n");
    return failures ? 1 : 0;
}
