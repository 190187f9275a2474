// grace/cuda/trace_sph.cuh -- SPH trace API (reference: cuda/trace_sph.cuh:22-241).
#pragma once
#include "grace/cuda/nodes.h"
#include "grace/generic/raydata.h"
#include "grace/ray.h"

namespace grace {

const static int N_table = 51;

// Line integrals of the Gadget-2 cubic spline kernel at impact parameter b/h = i/50.
template <typename Real>
struct KernelIntegrals {
    static const Real* table_ptr()
    {
        static Real t[N_table];
        static bool init = false;
        if (!init) {
            int n = 0;
            const double* src = grace_b200_kernel_integral_table(&n);
            for (int i = 0; i < N_table; ++i) t[i] = static_cast<Real>(src[i]);
            init = true;
        }
        return t;
    }
};

namespace detail {
inline const grace_b200_ray* rays_ptr(const Ray* r) { return reinterpret_cast<const grace_b200_ray*>(r); }

// GRACE_DEBUG builds of the reference assert inside the kernel when a traversal stack overflows
// (bintree_trace.cuh:162-164) and synchronise after every launch (error.h:57-64).  The B200
// kernels raise a device-side flag instead (1 = stack overflow, 2 = walk did not terminate: a
// malformed tree); with GRACE_DEBUG defined it is read back after every trace -- one stream
// synchronisation, as in the reference's debug builds -- and reported the way error.h reports.
inline void debug_check_trace(const char* what)
{
#ifdef GRACE_DEBUG
    int flag = 0;
    GRACE_B200_CHECK(grace_b200_device_error(context(), &flag, nullptr));
    if (flag != 0) {
        std::fprintf(stderr, "**** GRACE device-side error %d in %s (%s)\n", flag, what,
                     flag == 1 ? "traversal stack overflow" : "traversal did not terminate");
        std::exit(EXIT_FAILURE);
    }
#else
    (void)what;
#endif
}
}

// All throw std::invalid_argument unless d_rays.size() % 32 == 0.
template <typename RayVec, typename SphereVec, typename IntVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void trace_hitcounts_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree,
                                    IntVec& d_hit_counts)
{
    const grace_b200_tree t = detail::tree_view(d_tree);
    GRACE_B200_CHECK(grace_b200_trace_hitcounts_f4(
        detail::context(), detail::rays_ptr(detail::raw(d_rays.data())), d_rays.size(),
        reinterpret_cast<const float*>(detail::raw(d_spheres.data())), d_spheres.size(), &t,
        detail::raw(d_hit_counts.data()), nullptr));
    detail::debug_check_trace("trace_hitcounts_sph");
}

template <typename RayVec, typename SphereVec, typename RealVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void trace_cumulative_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree,
                                     RealVec& d_cumulated)
{
    const grace_b200_tree t = detail::tree_view(d_tree);
    GRACE_B200_CHECK(grace_b200_trace_cumulative_f4(
        detail::context(), detail::rays_ptr(detail::raw(d_rays.data())), d_rays.size(),
        reinterpret_cast<const float*>(detail::raw(d_spheres.data())), d_spheres.size(), &t,
        detail::raw(d_cumulated.data()), nullptr));
    detail::debug_check_trace("trace_cumulative_sph");
}

namespace detail {
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec>
inline void trace_lists(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, IntVec& d_ray_offsets,
                        IdxVec& d_hit_indices, RealVec& d_hit_integrals, RealVec& d_hit_distances, bool sentinels,
                        int index_sentinel, float integral_sentinel, float distance_sentinel)
{
    const grace_b200_tree t = tree_view(d_tree);
    const grace_b200_ray* rp = rays_ptr(raw(d_rays.data()));
    const float* sp = reinterpret_cast<const float*>(raw(d_spheres.data()));
    long long total = 0;
    GRACE_B200_CHECK(grace_b200_trace_hits_count_f4(context(), rp, d_rays.size(), sp, d_spheres.size(), &t,
                                                    sentinels ? 1 : 0, raw(d_ray_offsets.data()), &total, nullptr));
    if (sentinels) {
        d_hit_indices.resize((size_t)total, index_sentinel);
        d_hit_integrals.resize((size_t)total, integral_sentinel);
        d_hit_distances.resize((size_t)total, distance_sentinel);
    } else {
        d_hit_indices.resize((size_t)total);
        d_hit_integrals.resize((size_t)total);
        d_hit_distances.resize((size_t)total);
    }
    if (total > 0)
        GRACE_B200_CHECK(grace_b200_trace_hits_fill_f4(context(), rp, d_rays.size(), sp, d_spheres.size(), &t,
                                                       raw(d_ray_offsets.data()), raw(d_hit_indices.data()),
                                                       raw(d_hit_integrals.data()), raw(d_hit_distances.data()),
                                                       nullptr));
    debug_check_trace("trace_sph");
}
} // namespace detail

// d_ray_offsets must hold one int per ray; the three hit vectors are resized by the call.
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void trace_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree, IntVec& d_ray_offsets,
                          IdxVec& d_hit_indices, RealVec& d_hit_integrals, RealVec& d_hit_distances)
{
    detail::trace_lists(d_rays, d_spheres, d_tree, d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances,
                        false, 0, 0.f, 0.f);
}

// Each ray's segment ends with one slot holding the sentinels.
template <typename RayVec, typename SphereVec, typename IntVec, typename IdxVec, typename RealVec, typename Real, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST void trace_with_sentinels_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree,
                                         IntVec& d_ray_offsets, IdxVec& d_hit_indices, const int index_sentinel,
                                         RealVec& d_hit_integrals, const Real integral_sentinel,
                                         RealVec& d_hit_distances, const Real distance_sentinel)
{
    // Sentinel-filled resize of possibly non-empty vectors: start from empty like a fresh call.
    d_hit_indices.resize(0); d_hit_integrals.resize(0); d_hit_distances.resize(0);
    detail::trace_lists(d_rays, d_spheres, d_tree, d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances,
                        true, index_sentinel, (float)integral_sentinel, (float)distance_sentinel);
}

// Not in the reference: sorted hit lists of a ray set too large for one trace_sph call (a 4096^2
// projection of 2^24 particles has ~5e10 hits; trace_sph's offsets are int, cuda/trace_sph.cuh:117),
// streamed in ray tiles: trace_sph + sort_by_distance per tile in library-owned buffers of
// `hit_budget` hits, then
//     consume(size_t first_ray, size_t n_rays, const int* d_ray_offsets, long long n_hits,
//             const int* d_hit_indices, const float* d_hit_integrals, const float* d_hit_distances,
//             cudaStream_t stream)
// which must enqueue its work on `stream`; it overlaps the counting traversal of the next tile.
// Returns the total number of hits.
template <typename RayVec, typename SphereVec, typename Consume, detail::if_elem<SphereVec, float4> = 0>
GRACE_HOST long long trace_sorted_tiles_sph(const RayVec& d_rays, const SphereVec& d_spheres, const Tree& d_tree,
                                            const size_t hit_budget, Consume consume, const size_t rays_per_tile = 0)
{
    struct Tramp {
        static int call(void* user, size_t first_ray, size_t n_rays, const int* off, long long n_hits, const int* idx,
                        const float* integ, const float* dist, void* stream)
        {
            (*static_cast<Consume*>(user))(first_ray, n_rays, off, n_hits, idx, integ, dist, (cudaStream_t)stream);
            return 0;
        }
    };
    const grace_b200_tree t = detail::tree_view(d_tree);
    long long total = 0;
    GRACE_B200_CHECK(grace_b200_trace_sorted_tiles_f4(detail::context(), reinterpret_cast<const grace_b200_ray*>(detail::raw(d_rays.data())), d_rays.size(),
                                                      detail::f4(detail::raw(d_spheres.data())), d_spheres.size(), &t, hit_budget,
                                                      rays_per_tile, &Tramp::call, &consume, &total, nullptr));
    return total;
}

} // namespace grace

#ifdef __CUDACC__
#include "grace/cuda/sph_double.cuh"   // double4 spheres through the header templates
#endif
