// grace/cuda/util/extrema.cuh -- bounds of a set of float4 primitives (reference:
// cuda/util/extrema.cuh:189-230 min_max_x, :456-731 min_vec3/max_vec3/min_vec4/max_vec4),
// one fused reduction instead of one Thrust pass per call.
#pragma once
#include "grace/device_vector.h"

namespace grace {
namespace detail {
inline void minmax8(const float4* d_ptr, size_t n, float out[8])
{
    float* d_out = nullptr;
    GRACE_CUDA_CHECK(cudaMalloc((void**)&d_out, 8 * sizeof(float)));
    GRACE_B200_CHECK(grace_b200_minmax_f4(context(), (const float*)d_ptr, n, d_out, nullptr));
    GRACE_CUDA_CHECK(cudaMemcpy(out, d_out, 8 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_out);
}
} // namespace detail

inline void min_vec3(const float4* d_ptr, size_t n, float3* mins)
{ float v[8]; detail::minmax8(d_ptr, n, v); mins->x = v[0]; mins->y = v[1]; mins->z = v[2]; }
inline void max_vec3(const float4* d_ptr, size_t n, float3* maxs)
{ float v[8]; detail::minmax8(d_ptr, n, v); maxs->x = v[4]; maxs->y = v[5]; maxs->z = v[6]; }
inline void min_vec4(const float4* d_ptr, size_t n, float4* mins)
{ float v[8]; detail::minmax8(d_ptr, n, v); mins->x = v[0]; mins->y = v[1]; mins->z = v[2]; mins->w = v[3]; }
inline void max_vec4(const float4* d_ptr, size_t n, float4* maxs)
{ float v[8]; detail::minmax8(d_ptr, n, v); maxs->x = v[4]; maxs->y = v[5]; maxs->z = v[6]; maxs->w = v[7]; }
inline void min_max_x(const float4* d_ptr, size_t n, float* min_x, float* max_x)
{ float v[8]; detail::minmax8(d_ptr, n, v); *min_x = v[0]; *max_x = v[4]; }

template <typename Vec> inline void min_vec3(const Vec& d, float3* m) { min_vec3(detail::raw(d.data()), d.size(), m); }
template <typename Vec> inline void max_vec3(const Vec& d, float3* m) { max_vec3(detail::raw(d.data()), d.size(), m); }
template <typename Vec> inline void min_vec4(const Vec& d, float4* m) { min_vec4(detail::raw(d.data()), d.size(), m); }
template <typename Vec> inline void max_vec4(const Vec& d, float4* m) { max_vec4(detail::raw(d.data()), d.size(), m); }
template <typename Vec> inline void min_max_x(const Vec& d, float* a, float* b) { min_max_x(detail::raw(d.data()), d.size(), a, b); }

} // namespace grace
