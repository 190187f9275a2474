// grace/cuda/sph_double.cuh -- the SPH API for double-precision spheres, `double4 {x, y, z, h}`
// (reference: the same templates of cuda/build_sph.cuh:19-124 and cuda/trace_sph.cuh:58-241
// instantiated with Real4 = double4).  CUDA only; included by build_sph.cuh and
// trace_sph.cuh when compiled by nvcc.
//
// As in the reference, centroids and hence Morton keys, deltas and node boxes are single
// precision (CentroidSphere / AABBSphere return float3: the key of a double4 sphere is the key of
// its float-rounded centre, generic/functors/centroid.h:33-40), while intersection, kernel
// integrals and accumulation run in double.  These overloads are the reference's own recipe
// spelled with this repo's generic templates (kernels/{morton,albvh,bintree_trace}.cuh), so they
// inherit their parity with the reference's generic path; the float4 overloads keep using the
// hand-written kernels behind the C ABI.
#pragma once
#ifdef __CUDACC__
#include "grace/cuda/functors/trace.cuh"
#include "grace/cuda/kernels/albvh.cuh"
#include "grace/cuda/kernels/bintree_trace.cuh"
#include "grace/cuda/kernels/morton.cuh"
#include "grace/cuda/sort_by_key.cuh"
#include "grace/generic/functors/albvh.h"
#include "grace/generic/functors/centroid.h"
#include "grace/generic/raydata.h"

namespace grace {

// ---- build (cuda/build_sph.cuh:19-124) ----
template <typename SphereVec, typename KeyVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys_sph(const SphereVec& d_spheres, KeyVec& d_keys)
{
    morton_keys(d_spheres, d_keys, CentroidSphere());
}
template <typename Real3, typename SphereVec, typename KeyVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys_sph(const SphereVec& d_spheres, const Real3 bot, const Real3 top, KeyVec& d_keys)
{
    morton_keys(d_spheres, bot, top, d_keys, CentroidSphere());
}
namespace detail {
template <typename KeyType, typename SphereVec>
inline void sort_d4(SphereVec& d_spheres, const float3* bot, const float3* top)
{
    device_vector<KeyType> d_keys(d_spheres.size());
    if (bot) morton_keys(d_spheres, *bot, *top, d_keys, CentroidSphere());
    else morton_keys(d_spheres, d_keys, CentroidSphere());
    sort_by_key(d_keys, d_spheres);
}
template <typename Real3> inline float3 to_f3(const Real3 v) { return make_float3((float)v.x, (float)v.y, (float)v.z); }
} // namespace detail
template <typename SphereVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys30_sort_sph(SphereVec& d_spheres) { detail::sort_d4<uinteger32>(d_spheres, nullptr, nullptr); }
template <typename Real3, typename SphereVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys30_sort_sph(SphereVec& d_spheres, const Real3 bot, const Real3 top)
{
    const float3 b = detail::to_f3(bot), t = detail::to_f3(top);
    detail::sort_d4<uinteger32>(d_spheres, &b, &t);
}
template <typename SphereVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys63_sort_sph(SphereVec& d_spheres) { detail::sort_d4<uinteger64>(d_spheres, nullptr, nullptr); }
template <typename Real3, typename SphereVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void morton_keys63_sort_sph(SphereVec& d_spheres, const Real3 bot, const Real3 top)
{
    const float3 b = detail::to_f3(bot), t = detail::to_f3(top);
    detail::sort_d4<uinteger64>(d_spheres, &b, &t);
}
template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void euclidean_deltas_sph(const SphereVec& d_spheres, DeltaVec& d_deltas)
{
    compute_deltas(d_spheres, d_deltas, DeltaEuclidean<const double4*, CentroidSphere>());
}
template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void surface_area_deltas_sph(const SphereVec& d_spheres, DeltaVec& d_deltas)
{
    compute_deltas(d_spheres, d_deltas, DeltaSurfaceArea<const double4*, AABBSphere>());
}
template <typename SphereVec, typename DeltaVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void ALBVH_sph(const SphereVec& d_spheres, const DeltaVec& d_deltas, Tree& d_tree)
{
    build_ALBVH(d_tree, d_spheres, d_deltas, AABBSphere());
}

// ---- trace (cuda/trace_sph.cuh:58-241) ----
template <typename RayVec, typename SphereVec, typename IntVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void trace_hitcounts_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, IntVec& d_hit_counts)
{
    trace_texref<RayData_datum<int> >(d_rays, d_spheres, d_tree, 0, Init_null(), Intersect_sphere_bool(), OnHit_increment(),
                                      RayEntry_null(), RayExit_to_array<int>(detail::raw(d_hit_counts.data())));
}
namespace detail {
inline device_vector<double> kernel_table_on_device()
{
    int n = 0;
    const double* t = grace_b200_kernel_integral_table(&n);
    return device_vector<double>(std::vector<double>(t, t + n));
}
} // namespace detail
template <typename RayVec, typename SphereVec, typename RealVec, detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void trace_cumulative_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, RealVec& d_cumulated)
{
    typedef typename detail::elem_of<RealVec>::type Real;
    const device_vector<double> d_lookup = detail::kernel_table_on_device();
    trace_texref<RayData_sphere<Real, Real> >(d_rays, d_spheres, d_tree, sizeof(double) * d_lookup.size(),
                                              InitGlobalToSmem<double>(d_lookup.data(), (int)d_lookup.size()),
                                              Intersect_sphere_b2dist(), OnHit_sphere_cumulate((int)d_lookup.size()),
                                              RayEntry_null(), RayExit_to_array<Real>(detail::raw(d_cumulated.data())));
}
namespace detail {
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec>
inline void trace_lists_d4(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, IntVec& d_ray_offsets,
                           IdxVec& d_hit_indices, RealVec& d_hit_integrals, RealVec& d_hit_distances, bool sentinels,
                           typename elem_of<IdxVec>::type index_sentinel, typename elem_of<RealVec>::type integral_sentinel,
                           typename elem_of<RealVec>::type distance_sentinel)
{
    typedef typename elem_of<IdxVec>::type IndexType;
    typedef typename elem_of<RealVec>::type Real;
    trace_hitcounts_sph(d_rays, d_spheres, d_tree, d_ray_offsets);
    // offsets (+ ray index with sentinels, trace_sph.cuh:196-207) and the total, in one scan
    device_vector<long long> d_total(1);
    GRACE_B200_CHECK(grace_b200_exclusive_scan_i32(context(), raw(d_ray_offsets.data()), raw(d_ray_offsets.data()),
                                                   d_ray_offsets.size(), d_total.data(), nullptr));
    size_t total = (size_t)d_total.to_host()[0];
    if (sentinels) {
        const size_t n = d_ray_offsets.size();
        std::vector<int> off(n);
        GRACE_CUDA_CHECK(cudaMemcpy(off.data(), raw(d_ray_offsets.data()), n * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) off[i] += (int)i;
        GRACE_CUDA_CHECK(cudaMemcpy(raw(d_ray_offsets.data()), off.data(), n * sizeof(int), cudaMemcpyHostToDevice));
        total += n;
        d_hit_indices.resize(0); d_hit_integrals.resize(0); d_hit_distances.resize(0);
        d_hit_indices.resize(total, index_sentinel);
        d_hit_integrals.resize(total, integral_sentinel);
        d_hit_distances.resize(total, distance_sentinel);
    } else {
        d_hit_indices.resize(total); d_hit_integrals.resize(total); d_hit_distances.resize(total);
    }
    const device_vector<double> d_lookup = kernel_table_on_device();
    trace_texref<RayData_sphere<int, Real> >(d_rays, d_spheres, d_tree, sizeof(double) * d_lookup.size(),
        InitGlobalToSmem<double>(d_lookup.data(), (int)d_lookup.size()), Intersect_sphere_b2dist(),
        OnHit_sphere_individual<IndexType, Real>(raw(d_hit_indices.data()), raw(d_hit_integrals.data()),
                                                 raw(d_hit_distances.data()), (int)d_lookup.size()),
        RayEntry_from_array<int>(raw(d_ray_offsets.data())), RayExit_null());
}
} // namespace detail
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec,
          detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void trace_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, IntVec& d_ray_offsets,
                          IdxVec& d_hit_indices, RealVec& d_hit_integrals, RealVec& d_hit_distances)
{
    detail::trace_lists_d4(d_rays, d_spheres, d_tree, d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances,
                           false, 0, 0, 0);
}
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec, typename Real,
          detail::if_elem<SphereVec, double4> = 0>
GRACE_HOST void trace_with_sentinels_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree,
                                         IntVec& d_ray_offsets, IdxVec& d_hit_indices, const int index_sentinel,
                                         RealVec& d_hit_integrals, const Real integral_sentinel,
                                         RealVec& d_hit_distances, const Real distance_sentinel)
{
    detail::trace_lists_d4(d_rays, d_spheres, d_tree, d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances,
                           true, index_sentinel, integral_sentinel, distance_sentinel);
}

} // namespace grace
#endif // __CUDACC__
