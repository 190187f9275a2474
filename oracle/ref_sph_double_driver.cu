// ref_sph_double_driver.cu -- the SPH API on double4 spheres through the REFERENCE's headers
// (patched only for CUDA-12 compatibility, oracle/patch_ref.py), same source as
// tests/cpp/sph_double_test.cu (tests/cpp/sph_double_common.cuh).  TEST INFRASTRUCTURE.
#include <curand_kernel.h>

#include "grace/cuda/nodes.h"
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/trace_sph.cuh"
#include "grace/ray.h"

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

template <typename T> using DV = thrust::device_vector<T>;
template <typename T> static DV<T> to_device(const std::vector<T>& h) { return DV<T>(h.begin(), h.end()); }
template <typename T> static std::vector<T> to_host(const DV<T>& d)
{
    thrust::host_vector<T> h = d;
    return std::vector<T>(h.begin(), h.end());
}

#include "sph_double_common.cuh"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    return sph_double_run(strtoul(argv[1], 0, 10), strtoul(argv[2], 0, 10), atoi(argv[3]), argv[4]);
}
