// grace/cuda/device/intrinsics.cuh -- three-input integer min/max on float bit patterns
// (reference: include/grace/cuda/device/intrinsics.cuh:8-52, where they are PTX `vmin/vmax`
// video instructions that sm_100 expands into shift/select chains).  Here they are the DPX
// three-input min/max of sm_90+ (one VIMNMX3 each); same values for every input.
#pragma once

namespace grace {

// min(min(a, b), c)
__device__ __forceinline__ int min_vmin(int a, int b, int c) { return __vimin3_s32(a, b, c); }
// max(max(a, b), c)
__device__ __forceinline__ int max_vmax(int a, int b, int c) { return __vimax3_s32(a, b, c); }
// max(min(a, b), c)
__device__ __forceinline__ int max_vmin(int a, int b, int c) { return max(min(a, b), c); }
// min(max(a, b), c)
__device__ __forceinline__ int min_vmax(int a, int b, int c) { return min(max(a, b), c); }

__device__ __forceinline__ float minf_vminf(float f1, float f2, float f3)
{ return __int_as_float(min_vmin(__float_as_int(f1), __float_as_int(f2), __float_as_int(f3))); }
__device__ __forceinline__ float maxf_vmaxf(float f1, float f2, float f3)
{ return __int_as_float(max_vmax(__float_as_int(f1), __float_as_int(f2), __float_as_int(f3))); }
__device__ __forceinline__ float minf_vmaxf(float f1, float f2, float f3)
{ return __int_as_float(min_vmax(__float_as_int(f1), __float_as_int(f2), __float_as_int(f3))); }
__device__ __forceinline__ float maxf_vminf(float f1, float f2, float f3)
{ return __int_as_float(max_vmin(__float_as_int(f1), __float_as_int(f2), __float_as_int(f3))); }

} // namespace grace
