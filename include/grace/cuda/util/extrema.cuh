// grace/cuda/util/extrema.cuh -- component-wise bounds of a set of vectors (reference:
// cuda/util/extrema.cuh:189-455 min_max_{x,y,z,w}, :456-731 min_vec3/max_vec3/min_vec4/max_vec4).
// float4 data (the SPH primitives) go through the C ABI: one fused reduction over all four
// components with the result returned through the context's pinned scalars -- no allocation and
// one blocking read per call.  Any other element type with .x/.y/.z[/.w] members (the float3
// centroids of the tree-build profilers, double4) is reduced by the header template below.
#pragma once
#include "grace/device_vector.h"

#include <cfloat>

namespace grace {
namespace detail {

inline void minmax8(const float4* d_ptr, size_t n, float out[8])
{
    GRACE_B200_CHECK(grace_b200_minmax_f4_host(context(), (const float*)d_ptr, n, out, nullptr));
}

template <typename T> struct vec_traits;
template <> struct vec_traits<float3>  { typedef float  S; enum { N = 3 }; };
template <> struct vec_traits<float4>  { typedef float  S; enum { N = 4 }; };
template <> struct vec_traits<double3> { typedef double S; enum { N = 3 }; };
template <> struct vec_traits<double4> { typedef double S; enum { N = 4 }; };

#ifdef __CUDACC__
__host__ __device__ inline float comp_w(const float3&) { return 0.f; }
__host__ __device__ inline float comp_w(const float4& v) { return v.w; }
__host__ __device__ inline double comp_w(const double3&) { return 0.; }
__host__ __device__ inline double comp_w(const double4& v) { return v.w; }
template <typename V> __host__ __device__ inline typename vec_traits<V>::S comp_of(const V& v, int k)
{ return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : comp_w(v); }

// out[0..4) = component-wise minima, out[4..8) = maxima; one block per 4096 elements folds into
// `partial`, the last block to finish (ticket) folds the partials.
template <typename V, typename S>
__global__ void minmax_generic_kernel(const V* __restrict__ data, size_t n, S* partial, unsigned* ticket, S* out)
{
    __shared__ S s_lo[4][8], s_hi[4][8];
    __shared__ bool s_last;
    S lo[4], hi[4];
    for (int k = 0; k < 4; ++k) { lo[k] = (S)FLT_MAX * (S)FLT_MAX; hi[k] = -lo[k]; }       // +inf / -inf in either precision
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const V v = data[i];
        for (int k = 0; k < vec_traits<V>::N; ++k) { const S c = comp_of(v, k); lo[k] = c < lo[k] ? c : lo[k]; hi[k] = c > hi[k] ? c : hi[k]; }
    }
    auto fold = [&]() {
        for (int k = 0; k < 4; ++k)
            for (int o = 16; o > 0; o >>= 1) {
                const S a = __shfl_xor_sync(0xffffffffu, lo[k], o), b = __shfl_xor_sync(0xffffffffu, hi[k], o);
                lo[k] = a < lo[k] ? a : lo[k]; hi[k] = b > hi[k] ? b : hi[k];
            }
    };
    fold();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int k = 0; k < 4; ++k) { s_lo[k][warp] = lo[k]; s_hi[k][warp] = hi[k]; }
    __syncthreads();
    if (warp == 0) {
        for (int k = 0; k < 4; ++k) { lo[k] = lane < (int)(blockDim.x >> 5) ? s_lo[k][lane] : lo[k]; hi[k] = lane < (int)(blockDim.x >> 5) ? s_hi[k][lane] : hi[k]; }
        fold();
        if (lane == 0) {
            for (int k = 0; k < 4; ++k) { partial[8 * blockIdx.x + k] = lo[k]; partial[8 * blockIdx.x + 4 + k] = hi[k]; }
            __threadfence();
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    for (int k = 0; k < 4; ++k) { lo[k] = (S)FLT_MAX * (S)FLT_MAX; hi[k] = -lo[k]; }
    for (unsigned b = lane; b < gridDim.x; b += 32)
        for (int k = 0; k < 4; ++k) {
            const S a = ((volatile S*)partial)[8 * b + k], c = ((volatile S*)partial)[8 * b + 4 + k];
            lo[k] = a < lo[k] ? a : lo[k]; hi[k] = c > hi[k] ? c : hi[k];
        }
    fold();
    if (lane == 0) { for (int k = 0; k < 4; ++k) { out[k] = lo[k]; out[4 + k] = hi[k]; } *ticket = 0u; }
}

template <typename V>
inline void minmax8(const V* d_ptr, size_t n, typename vec_traits<V>::S out[8])
{
    typedef typename vec_traits<V>::S S;
    const int blocks = (int)((n + 4095) / 4096 < 1184 ? (n + 4095) / 4096 : 1184) + (n == 0);
    device_vector<S> scratch(8 * (size_t)blocks + 8 + 2);
    S* partial = scratch.data();
    S* d_out = partial + 8 * (size_t)blocks;
    unsigned* ticket = (unsigned*)(d_out + 8);
    GRACE_CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(unsigned)));
    minmax_generic_kernel<V, S><<<blocks, 256>>>(d_ptr, n, partial, ticket, d_out);
    GRACE_KERNEL_CHECK();
    GRACE_CUDA_CHECK(cudaMemcpy(out, d_out, 8 * sizeof(S), cudaMemcpyDeviceToHost));
}

#endif // __CUDACC__ (element types other than float4 need the header kernel, hence nvcc)

} // namespace detail

// ---- pointer forms (d_ptr is DEVICE memory), any vector type with .x/.y/.z[/.w] ----
#define GRACE_B200_EXTREMA_BODY(V) typename detail::vec_traits<V>::S v[8]; detail::minmax8(d_ptr, n, v);
template <typename V, typename V3> inline void min_vec3(const V* d_ptr, size_t n, V3* mins)
{ GRACE_B200_EXTREMA_BODY(V) mins->x = v[0]; mins->y = v[1]; mins->z = v[2]; }
template <typename V, typename V3> inline void max_vec3(const V* d_ptr, size_t n, V3* maxs)
{ GRACE_B200_EXTREMA_BODY(V) maxs->x = v[4]; maxs->y = v[5]; maxs->z = v[6]; }
template <typename V, typename V4> inline void min_vec4(const V* d_ptr, size_t n, V4* mins)
{ GRACE_B200_EXTREMA_BODY(V) mins->x = v[0]; mins->y = v[1]; mins->z = v[2]; mins->w = v[3]; }
template <typename V, typename V4> inline void max_vec4(const V* d_ptr, size_t n, V4* maxs)
{ GRACE_B200_EXTREMA_BODY(V) maxs->x = v[4]; maxs->y = v[5]; maxs->z = v[6]; maxs->w = v[7]; }
template <typename V, typename T> inline void min_max_x(const V* d_ptr, size_t n, T* lo, T* hi)
{ GRACE_B200_EXTREMA_BODY(V) *lo = v[0]; *hi = v[4]; }
template <typename V, typename T> inline void min_max_y(const V* d_ptr, size_t n, T* lo, T* hi)
{ GRACE_B200_EXTREMA_BODY(V) *lo = v[1]; *hi = v[5]; }
template <typename V, typename T> inline void min_max_z(const V* d_ptr, size_t n, T* lo, T* hi)
{ GRACE_B200_EXTREMA_BODY(V) *lo = v[2]; *hi = v[6]; }
template <typename V, typename T> inline void min_max_w(const V* d_ptr, size_t n, T* lo, T* hi)
{ GRACE_B200_EXTREMA_BODY(V) *lo = v[3]; *hi = v[7]; }
#undef GRACE_B200_EXTREMA_BODY

// ---- container forms (grace::device_vector or thrust::device_vector) ----
#define GRACE_B200_EXTREMA_VEC(name, OutT)                                                                \
    template <typename Vec, typename OutT, typename = decltype(std::declval<const Vec&>().size())>          \
    inline void name(const Vec& d, OutT* a) { name(detail::raw(d.data()), d.size(), a); }
GRACE_B200_EXTREMA_VEC(min_vec3, V3)
GRACE_B200_EXTREMA_VEC(max_vec3, V3)
GRACE_B200_EXTREMA_VEC(min_vec4, V4)
GRACE_B200_EXTREMA_VEC(max_vec4, V4)
#undef GRACE_B200_EXTREMA_VEC
#define GRACE_B200_EXTREMA_VEC2(name)                                                                     \
    template <typename Vec, typename T, typename = decltype(std::declval<const Vec&>().size())>             \
    inline void name(const Vec& d, T* a, T* b) { name(detail::raw(d.data()), d.size(), a, b); }
GRACE_B200_EXTREMA_VEC2(min_max_x)
GRACE_B200_EXTREMA_VEC2(min_max_y)
GRACE_B200_EXTREMA_VEC2(min_max_z)
GRACE_B200_EXTREMA_VEC2(min_max_w)
#undef GRACE_B200_EXTREMA_VEC2

} // namespace grace
