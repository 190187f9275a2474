// ref_gadget_driver.cu -- loads a Gadget-2 file with the REFERENCE's own reader
// (tests/helper/read_gadget.cuh:69-167, unmodified) into a device_vector, times it and dumps
// the float4 records.  TEST INFRASTRUCTURE, built by oracle/build_ref.sh into oracle/_ref/.
//   ref_gadget_driver <file> <out.bin>
#include "helper/read_gadget.cuh"

#include <chrono>
#include <cstdio>

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    cudaFree(0);
    thrust::device_vector<float4> d_pos;
    const auto t0 = std::chrono::steady_clock::now();
    read_gadget(argv[1], d_pos);
    cudaDeviceSynchronize();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    thrust::host_vector<float4> h = d_pos;
    FILE* f = fopen(argv[2], "wb");
    if (!f) return 2;
    fwrite(thrust::raw_pointer_cast(h.data()), sizeof(float4), h.size(), f);
    fclose(f);
    printf("{\"n_gas\": %zu, \"seconds_file_to_device\": %.6f}\n", h.size(), s);
    return 0;
}
