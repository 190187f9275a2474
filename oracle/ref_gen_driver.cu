// ref_gen_driver.cu -- runs every ray generator of the REFERENCE (GRACE, patched only for
// CUDA-12 compatibility by oracle/patch_ref.py) through its public API with the parameters
// given on the command line and dumps the rays.
//
// TEST INFRASTRUCTURE: compiled by oracle/build_ref.sh into oracle/_ref/ref_gen_driver from the
// sources under /root/reference (never copied into this repository).  Pins the product's
// generators bit-for-bit (tests/test_gpu_vs_reference.py, tests/golden/).
//
//   ref_gen_driver <outdir> <points.bin> <n_random> <seed> <res_x> <res_y>
// points.bin: P x float3 end points for one_to_many_rays.  Outputs (R x 7 floats each):
//   octant.bin (MPM), o2m_nosort.bin, o2m_dirsort.bin, o2m_endsort_aabb.bin,
//   plane_parallel.bin, ortho.bin, pinhole.bin
#include <curand_kernel.h>

#include "grace/cuda/gen_rays.cuh"
#include "grace/ray.h"
#include "grace/types.h"

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static void dump(const std::string& dir, const char* name, const thrust::device_vector<grace::Ray>& d)
{
    thrust::host_vector<grace::Ray> h = d;
    FILE* f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror(name); exit(2); }
    fwrite(thrust::raw_pointer_cast(h.data()), sizeof(grace::Ray), h.size(), f);
    fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: see header\n"); return 2; }
    const std::string out = argv[1];
    const size_t n_random = strtoull(argv[3], NULL, 10);
    const unsigned long long seed = strtoull(argv[4], NULL, 10);
    const int rx = atoi(argv[5]), ry = atoi(argv[6]);

    std::vector<float3> h_pts;
    {
        FILE* f = fopen(argv[2], "rb");
        if (!f) { perror(argv[2]); return 2; }
        fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
        h_pts.resize(bytes / sizeof(float3));
        if (fread(h_pts.data(), sizeof(float3), h_pts.size(), f) != h_pts.size()) return 2;
        fclose(f);
    }
    thrust::device_vector<float3> d_pts(h_pts.begin(), h_pts.end());
    thrust::device_vector<grace::Ray> d_rays;

    d_rays.resize(n_random);
    grace::uniform_random_rays_single_octant(d_rays, 0.5f, 0.25f, 0.125f, 2.0f, grace::MPM, seed);
    dump(out, "octant.bin", d_rays);

    d_rays.resize(h_pts.size());
    grace::one_to_many_rays(d_rays, 0.1f, 0.2f, 0.3f, d_pts, grace::NoSort);
    dump(out, "o2m_nosort.bin", d_rays);
    grace::one_to_many_rays(d_rays, 0.1f, 0.2f, 0.3f, d_pts, grace::DirectionSort);
    dump(out, "o2m_dirsort.bin", d_rays);
    grace::one_to_many_rays(d_rays, 0.1f, 0.2f, 0.3f, d_pts, make_float3(0.f, 0.f, 0.f), make_float3(1.f, 1.f, 1.f));
    dump(out, "o2m_endsort_aabb.bin", d_rays);

    d_rays.resize((size_t)rx * ry);
    grace::plane_parallel_random_rays(d_rays, rx, ry, make_float3(-0.1f, -0.2f, 1.5f), make_float3(1.3f, 0.f, 0.f),
                                      make_float3(0.f, 1.1f, 0.f), 3.0f, seed);
    dump(out, "plane_parallel.bin", d_rays);
    grace::orthographic_projection_rays(d_rays, rx, ry, make_float3(0.5f, 0.4f, 2.0f), make_float3(0.45f, 0.5f, 0.5f),
                                        make_float3(0.f, 1.f, 0.1f), 1.25f, 4.0f);
    dump(out, "ortho.bin", d_rays);
    grace::pinhole_camera_rays(d_rays, rx, ry, make_float3(0.5f, 0.4f, 2.0f), make_float3(0.45f, 0.5f, 0.5f),
                               make_float3(0.f, 1.f, 0.1f), 0.9f, 4.0f);
    dump(out, "pinhole.bin", d_rays);
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e)); return 3; }
    printf("{\"ok\": true}\n");
    return 0;
}
