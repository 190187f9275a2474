// grace/cuda/kernels/morton.cuh -- Morton keys of arbitrary primitives through a user centroid
// functor (reference: cuda/kernels/morton.cuh:30-189).  CUDA only.  The functor
// is evaluated by a templated kernel here; bounds and keys come from the library (C ABI) on the
// centroid array, in the reference's arithmetic: q = (Key)(scale * (c - min)), scale = span / (top - bot).
#pragma once
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/types.h"

#include <climits>
#include <iterator>

namespace grace {

namespace morton {
template <typename TPrimitive, typename CentroidFunc>
__global__ void centroids_kernel(const TPrimitive* __restrict__ prims, const size_t n, float4* __restrict__ centroids,
                                 const CentroidFunc centroid)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float3 c = centroid(prims[i]);
        centroids[i] = make_float4(c.x, c.y, c.z, 0.f);
    }
}
inline int keys_from_centroids(const float4* c, size_t n, const float* b6, uinteger32* keys)
{
    return grace_b200_morton_keys30_f4(detail::context(), (const float*)c, n, b6, keys, nullptr);
}
inline int keys_from_centroids(const float4* c, size_t n, const float* b6, uinteger64* keys)
{
    return grace_b200_morton_keys63_f4(detail::context(), (const float*)c, n, b6, keys, nullptr);
}
} // namespace morton

template <typename TPrimitive, typename Real3, typename KeyType, typename CentroidFunc>
GRACE_HOST void morton_keys(const TPrimitive* d_prims, const size_t N_primitives, const Real3 AABB_bot, const Real3 AABB_top,
                            KeyType* d_keys, const CentroidFunc centroid)
{
    if (N_primitives == 0) return;
    device_vector<float4> d_centroids(N_primitives);
    const int blocks = (int)((N_primitives + 255) / 256 < 4096 ? (N_primitives + 255) / 256 : 4096);
    morton::centroids_kernel<<<blocks, 256>>>(d_prims, N_primitives, d_centroids.data(), centroid);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    device_vector<float> d_b6(std::vector<float>{ (float)AABB_bot.x, (float)AABB_bot.y, (float)AABB_bot.z,
                                                  (float)AABB_top.x, (float)AABB_top.y, (float)AABB_top.z });
    GRACE_B200_CHECK(morton::keys_from_centroids(d_centroids.data(), N_primitives, d_b6.data(), d_keys));
    GRACE_CUDA_CHECK(cudaDeviceSynchronize());      // the temporaries die here
}

// As above, computing the bounds of the centroids first (optionally returned).
template <typename TPrimitive, typename KeyType, typename CentroidFunc>
GRACE_HOST void morton_keys(const TPrimitive* d_prims, const size_t N_primitives, KeyType* d_keys,
                            const CentroidFunc centroid, float3* const bots = NULL, float3* const tops = NULL)
{
    if (N_primitives == 0) return;
    device_vector<float4> d_centroids(N_primitives);
    const int blocks = (int)((N_primitives + 255) / 256 < 4096 ? (N_primitives + 255) / 256 : 4096);
    morton::centroids_kernel<<<blocks, 256>>>(d_prims, N_primitives, d_centroids.data(), centroid);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    device_vector<float> d_b6(6);
    GRACE_B200_CHECK(grace_b200_bounds_f4(detail::context(), (const float*)d_centroids.data(), N_primitives, d_b6.data(), nullptr));
    GRACE_B200_CHECK(morton::keys_from_centroids(d_centroids.data(), N_primitives, d_b6.data(), d_keys));
    const std::vector<float> b = d_b6.to_host();    // synchronises
    if (bots) { bots->x = b[0]; bots->y = b[1]; bots->z = b[2]; }
    if (tops) { tops->x = b[3]; tops->y = b[4]; tops->z = b[5]; }
}

template <typename PrimVec, typename Real3, typename KeyVec, typename CentroidFunc>
GRACE_HOST void morton_keys(const PrimVec& d_primitives, const Real3 AABB_bot, const Real3 AABB_top, KeyVec& d_keys,
                            const CentroidFunc centroid)
{
    morton_keys(detail::raw(d_primitives.data()), d_primitives.size(), AABB_bot, AABB_top, detail::raw(d_keys.data()), centroid);
}

template <typename PrimVec, typename KeyVec, typename CentroidFunc>
GRACE_HOST void morton_keys(const PrimVec& d_primitives, KeyVec& d_keys, const CentroidFunc centroid,
                            float3* const bots = NULL, float3* const tops = NULL)
{
    morton_keys(detail::raw(d_primitives.data()), d_primitives.size(), detail::raw(d_keys.data()), centroid, bots, tops);
}

} // namespace grace
