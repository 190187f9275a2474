// grace/cuda/build_sph.cuh -- SPH tree-build API (reference: cuda/build_sph.cuh:19-124),
// float4 spheres {x, y, z, h}.  Containers are any type with size()/data()/resize()
// (grace::device_vector or thrust::device_vector).
#pragma once
#include "grace/cuda/nodes.h"
#include "grace/generic/functors/albvh.h"
#include "grace/generic/functors/centroid.h"

namespace grace {
namespace detail {
inline const float* f4(const float4* p) { return reinterpret_cast<const float*>(p); }
inline float* f4(float4* p) { return reinterpret_cast<float*>(p); }
inline void keys(const float4* s, size_t n, const float* d_bounds6, uinteger32* k)
{ GRACE_B200_CHECK(grace_b200_morton_keys30_f4(context(), f4(s), n, d_bounds6, k, nullptr)); }
inline void keys(const float4* s, size_t n, const float* d_bounds6, uinteger64* k)
{ GRACE_B200_CHECK(grace_b200_morton_keys63_f4(context(), f4(s), n, d_bounds6, k, nullptr)); }
} // namespace detail

// Keys from the bounds of the sphere centres.
template <typename SphereVec, typename KeyVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys_sph(const SphereVec& d_spheres, KeyVec& d_keys)
{
    float* d_b = nullptr;
    GRACE_CUDA_CHECK(cudaMalloc((void**)&d_b, 6 * sizeof(float)));
    GRACE_B200_CHECK(grace_b200_bounds_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                          d_spheres.size(), d_b, nullptr));
    detail::keys(detail::raw(d_spheres.data()), d_spheres.size(), d_b, detail::raw(d_keys.data()));
    GRACE_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(d_b);
}

// Keys from explicit bounds.
template <typename Real3, typename SphereVec, typename KeyVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys_sph(const SphereVec& d_spheres, const Real3 bot, const Real3 top, KeyVec& d_keys)
{
    const float h[6] = { (float)bot.x, (float)bot.y, (float)bot.z, (float)top.x, (float)top.y, (float)top.z };
    float* d_b = nullptr;
    GRACE_CUDA_CHECK(cudaMalloc((void**)&d_b, sizeof(h)));
    GRACE_CUDA_CHECK(cudaMemcpy(d_b, h, sizeof(h), cudaMemcpyHostToDevice));
    detail::keys(detail::raw(d_spheres.data()), d_spheres.size(), d_b, detail::raw(d_keys.data()));
    GRACE_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(d_b);
}

template <typename SphereVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys30_sort_sph(SphereVec& d_spheres)
{
    GRACE_B200_CHECK(grace_b200_morton_sort_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                               d_spheres.size(), 30, nullptr, nullptr, nullptr, nullptr));
}
template <typename Real3, typename SphereVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys30_sort_sph(SphereVec& d_spheres, const Real3 bot, const Real3 top)
{
    const float b[3] = { (float)bot.x, (float)bot.y, (float)bot.z }, t[3] = { (float)top.x, (float)top.y, (float)top.z };
    GRACE_B200_CHECK(grace_b200_morton_sort_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                               d_spheres.size(), 30, b, t, nullptr, nullptr));
}
template <typename SphereVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys63_sort_sph(SphereVec& d_spheres)
{
    GRACE_B200_CHECK(grace_b200_morton_sort_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                               d_spheres.size(), 63, nullptr, nullptr, nullptr, nullptr));
}
template <typename Real3, typename SphereVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void morton_keys63_sort_sph(SphereVec& d_spheres, const Real3 bot, const Real3 top)
{
    const float b[3] = { (float)bot.x, (float)bot.y, (float)bot.z }, t[3] = { (float)top.x, (float)top.y, (float)top.z };
    GRACE_B200_CHECK(grace_b200_morton_sort_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                               d_spheres.size(), 63, b, t, nullptr, nullptr));
}

template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void euclidean_deltas_sph(const SphereVec& d_spheres, DeltaVec& d_deltas)
{
    GRACE_B200_CHECK(grace_b200_deltas_euclid_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                                 d_spheres.size(), detail::raw(d_deltas.data()), nullptr));
}
template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void surface_area_deltas_sph(const SphereVec& d_spheres, DeltaVec& d_deltas)
{
    GRACE_B200_CHECK(grace_b200_deltas_sarea_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())),
                                                d_spheres.size(), detail::raw(d_deltas.data()), nullptr));
}
namespace detail {
inline int xor_deltas(const uinteger32* k, size_t n, uinteger32* d) { return grace_b200_deltas_xor32(context(), k, n, d, nullptr); }
inline int xor_deltas(const uinteger64* k, size_t n, uinteger64* d) { return grace_b200_deltas_xor64(context(), k, n, d, nullptr); }
inline int delta_type(const float*) { return GRACE_B200_DELTA_F32; }
inline int delta_type(const uinteger32*) { return GRACE_B200_DELTA_U32; }
inline int delta_type(const uinteger64*) { return GRACE_B200_DELTA_U64; }
} // namespace detail
template <typename KeyVec, typename DeltaVec>
GRACE_HOST void XOR_deltas_sph(const KeyVec& d_keys, DeltaVec& d_deltas)
{
    GRACE_B200_CHECK(detail::xor_deltas(detail::raw(d_keys.data()), d_keys.size(), detail::raw(d_deltas.data())));
}

// Throws std::invalid_argument if the number of spheres is <= tree.max_per_leaf.
template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void ALBVH_sph(const SphereVec& d_spheres, const DeltaVec& d_deltas, Tree& d_tree)
{
    int L = 0;
    const auto* dp = detail::raw(d_deltas.data());
    const size_t n = d_spheres.size();
    if (d_tree.leaves.size() < n) d_tree.leaves.resize(n);
    // leaves first: their number sizes the node array (remove_empty_leaves, albvh.cuh:842-845), so a Tree
    // whose storage has not been touched yet never allocates the 4 (N - 1) int4 it was constructed for
    GRACE_B200_CHECK(grace_b200_albvh_leaves(detail::context(), dp, detail::delta_type(dp), n, d_tree.max_per_leaf,
                                             d_tree.leaves.data(), &L, nullptr));
    d_tree.nodes.resize(4 * (size_t)(L - 1));
    d_tree.leaves.resize((size_t)L);
    // the leaf-level deltas are still in the library's workspace from the leaves stage
    GRACE_B200_CHECK(grace_b200_albvh_nodes_f4(detail::context(), detail::f4(detail::raw(d_spheres.data())), d_tree.leaves.data(),
                                               (size_t)L, nullptr, detail::delta_type(dp), d_tree.nodes.data(),
                                               d_tree.root_index_ptr, nullptr));
}

} // namespace grace

#ifdef __CUDACC__
#include "grace/cuda/sph_double.cuh"   // double4 spheres through the header templates
#endif
