// sph_double_common.cuh -- the SPH recipe on double4 spheres, written ONLY against the API GRACE
// and this repo share, and included by tests/cpp/sph_double_test.cu (this repo's headers,
// grace::device_vector) and oracle/ref_sph_double_driver.cu (the reference's headers,
// thrust::device_vector).  The including file defines the alias template DV<T> and the helpers
// to_device(std::vector<T>) / to_host(DV<T>).  Test infrastructure.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static inline unsigned sd_hash(unsigned a)
{
    a = (a + 0x7ed55d16) + (a << 12); a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);  a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);  a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
// 53-bit mantissas so the double precision matters
static inline double sd_u01(unsigned i, unsigned k)
{
    const unsigned long long hi = sd_hash(i * 8u + k), lo = sd_hash(i * 8u + k + 0x9e3779b9u);
    return (double)(((hi << 21) ^ lo) & ((1ull << 53) - 1)) / 9007199254740992.0;
}

template <typename T>
static void sd_dump(const std::string& dir, const char* name, const std::vector<T>& v)
{
    FILE* f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror(name); exit(2); }
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

static int sph_double_run(size_t N, size_t R, int key_bits, const std::string& out)
{
    std::vector<double4> h_s(N);
    for (size_t i = 0; i < N; ++i)
        h_s[i] = make_double4(sd_u01((unsigned)i, 0), sd_u01((unsigned)i, 1), sd_u01((unsigned)i, 2), 0.002 + 0.03 * sd_u01((unsigned)i, 3));
    std::vector<grace::Ray> h_r(R);
    for (size_t i = 0; i < R; ++i) {
        const unsigned u = (unsigned)i + 4242u;
        double dx = 2 * sd_u01(u, 0) - 1, dy = 2 * sd_u01(u, 1) - 1, dz = 2 * sd_u01(u, 2) - 1;
        const double inv = 1.0 / sqrt(dx * dx + dy * dy + dz * dz + 1e-30);
        h_r[i].dx = (float)(dx * inv); h_r[i].dy = (float)(dy * inv); h_r[i].dz = (float)(dz * inv);
        h_r[i].ox = 0.5f; h_r[i].oy = 0.45f; h_r[i].oz = 0.55f; h_r[i].length = 2.f;
    }
    DV<double4> d_s = to_device(h_s);
    DV<grace::Ray> d_r = to_device(h_r);

    if (key_bits == 30) grace::morton_keys30_sort_sph(d_s); else grace::morton_keys63_sort_sph(d_s);
    DV<grace::uinteger64> d_keys(N);
    grace::morton_keys_sph(d_s, make_float3(0.f, 0.f, 0.f), make_float3(1.f, 1.f, 1.f), d_keys);
    DV<float> d_deltas(N + 1), d_sa(N + 1);
    grace::euclidean_deltas_sph(d_s, d_deltas);
    grace::surface_area_deltas_sph(d_s, d_sa);
    grace::Tree d_tree(N, 16);
    grace::ALBVH_sph(d_s, d_deltas, d_tree);

    DV<int> d_counts(R), d_offsets(R), d_idx;
    DV<double> d_cum(R), d_integ, d_dist;
    grace::trace_hitcounts_sph(d_r, d_s, d_tree, d_counts);
    grace::trace_cumulative_sph(d_r, d_s, d_tree, d_cum);
    grace::trace_sph(d_r, d_s, d_tree, d_offsets, d_idx, d_integ, d_dist);
    cudaDeviceSynchronize();

    sd_dump(out, "spheres.bin", to_host(d_s));
    sd_dump(out, "keys63.bin", to_host(d_keys));
    sd_dump(out, "deltas.bin", to_host(d_deltas));
    sd_dump(out, "sarea.bin", to_host(d_sa));
    std::vector<int4> leaves = to_host(d_tree.leaves), nodes = to_host(d_tree.nodes);
    nodes.resize(4 * (leaves.size() - 1));
    sd_dump(out, "leaves.bin", leaves);
    sd_dump(out, "nodes.bin", nodes);
    sd_dump(out, "counts.bin", to_host(d_counts));
    sd_dump(out, "cum.bin", to_host(d_cum));
    sd_dump(out, "offsets.bin", to_host(d_offsets));
    sd_dump(out, "hit_idx.bin", to_host(d_idx));
    sd_dump(out, "hit_integ.bin", to_host(d_integ));
    sd_dump(out, "hit_dist.bin", to_host(d_dist));
    long long hits = 0;
    for (int c : to_host(d_counts)) hits += c;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e)); return 3; }
    printf("{\"n\": %zu, \"rays\": %zu, \"key_bits\": %d, \"n_leaves\": %zu, \"hits\": %lld}\n", N, R, key_bits, leaves.size(), hits);
    return hits > 0 ? 0 : 1;
}
