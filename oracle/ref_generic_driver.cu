// ref_generic_driver.cu -- the REFERENCE's generic (user primitive + functor) path, GRACE's headers
// patched only for CUDA-12 compatibility (oracle/patch_ref.py), driven with the SAME user
// primitive and functors as tests/cpp/generic_test.cu (tests/cpp/generic_prims.cuh), dumping
// the same arrays.  TEST INFRASTRUCTURE, built by oracle/build_ref.sh into oracle/_ref/.
//   ref_generic_driver <n_tris> <n_rays> <outdir>
#include <curand_kernel.h>

#include "grace/cuda/nodes.h"
#include "grace/cuda/functors/trace.cuh"
#include "grace/cuda/kernels/albvh.cuh"
#include "grace/cuda/kernels/bintree_trace.cuh"
#include "grace/cuda/kernels/morton.cuh"
#include "grace/generic/functors/albvh.h"
#include "grace/ray.h"
#include "grace/types.h"

#include "generic_prims.cuh"

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <thrust/sort.h>

#include <cstdio>
#include <cstdlib>
#include <string>

template <typename T>
static void dump(const std::string& dir, const char* name, const thrust::device_vector<T>& d, size_t count)
{
    thrust::host_vector<T> h(d.begin(), d.begin() + count);
    FILE* f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror(name); exit(2); }
    fwrite(thrust::raw_pointer_cast(h.data()), sizeof(T), h.size(), f);
    fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    const size_t N = strtoul(argv[1], 0, 10), R = strtoul(argv[2], 0, 10);
    const std::string out = argv[3];
    const int max_per_leaf = 8;
    const std::vector<Tri> h_tris = make_tris(N, 0.02f);
    const std::vector<grace::Ray> h_rays = make_rays(R);

    thrust::device_vector<Tri> d_tris(h_tris.begin(), h_tris.end());
    thrust::device_vector<grace::uinteger32> d_keys(N), d_deltas(N + 1);
    float3 bot, top;
    grace::morton_keys(d_tris, d_keys, TriCentroid(), &bot, &top);
    thrust::sort_by_key(d_keys.begin(), d_keys.end(), d_tris.begin());
    grace::compute_deltas(d_keys, d_deltas, grace::DeltaXOR());
    grace::Tree d_tree(N, max_per_leaf);
    grace::build_ALBVH(d_tree, d_tris, d_deltas, TriAABB());

    thrust::device_vector<grace::Ray> d_rays(h_rays.begin(), h_rays.end());
    thrust::device_vector<int> d_closest(R), d_counts(R);
    grace::trace_texref<RayData_tri>(d_rays, d_tris, d_tree, 0, grace::Init_null(), RayIntersect_tri(), OnHit_tri(),
                                     RayEntry_tri(),
                                     grace::RayExit_to_array<int>(thrust::raw_pointer_cast(d_closest.data())));
    grace::trace<RayData_cnt>(d_rays, d_tris, d_tree, 0, grace::Init_null(), RayIntersect_any(), grace::OnHit_increment(),
                              grace::RayEntry_null(),
                              grace::RayExit_to_array<int>(thrust::raw_pointer_cast(d_counts.data())));
    cudaDeviceSynchronize();
    const size_t L = d_tree.leaves.size();
    dump(out, "keys.bin", d_keys, N);
    dump(out, "tris.bin", d_tris, N);
    dump(out, "deltas.bin", d_deltas, N + 1);
    dump(out, "leaves.bin", d_tree.leaves, L);
    dump(out, "nodes.bin", d_tree.nodes, 4 * (L - 1));
    dump(out, "closest.bin", d_closest, R);
    dump(out, "counts.bin", d_counts, R);
    int root = 0;
    cudaMemcpy(&root, d_tree.root_index_ptr, sizeof(int), cudaMemcpyDeviceToHost);
    FILE* f = fopen((out + "/root.bin").c_str(), "wb");
    fwrite(&root, 4, 1, f);
    fclose(f);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e)); return 3; }
    printf("{\"n_leaves\": %zu, \"root\": %d}\n", L, root);
    return 0;
}
