// sph_double_test.cu -- the SPH API on double4 spheres through this repo's headers.
//   sph_double_test <n> <rays> <key_bits> <outdir>
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/trace_sph.cuh"

template <typename T> using DV = grace::device_vector<T>;
template <typename T> static DV<T> to_device(const std::vector<T>& h) { return DV<T>(h); }
template <typename T> static std::vector<T> to_host(const DV<T>& d) { return d.to_host(); }

#include "sph_double_common.cuh"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    return sph_double_run(strtoul(argv[1], 0, 10), strtoul(argv[2], 0, 10), atoi(argv[3]), argv[4]);
}
