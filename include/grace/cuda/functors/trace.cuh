// grace/cuda/functors/trace.cuh -- the functors of the generic traversal (reference:
// cuda/functors/trace.cuh:18-235).  Signatures:
//   Init(BoundIter<char>)                                                  once per block
//   RayEntry / RayExit(int ray_idx, const Ray&, RayData&, BoundIter<char>)
//   Intersection(const Ray&, const TPrim&, RayData&, int lane, BoundIter<char>) -> bool
//   OnHit(int ray_idx, const Ray&, RayData&, int prim_idx, const TPrim&, int lane, BoundIter<char>)
// CUDA only: include from a .cu translation unit.
#pragma once
#include "grace/cuda/util/bound_iter.cuh"
#include "grace/generic/interpolate.h"
#include "grace/generic/intersect.h"
#include "grace/generic/meta.h"
#include "grace/ray.h"
#include "grace/types.h"

namespace grace {

namespace detail {

typedef gpu::BoundIter<char> SmemBlock;   // the user's shared-memory block as every functor receives it

// One hit's contribution to a column density: the kernel line integral at impact parameter
// sqrt(b2), for a particle of smoothing length h.  The table (n_table doubles) is read through
// the bounds-checked iterator.  Operation order is the reference's (cuda/functors/trace.cuh:183-191
// and :221-228): 1/h, then (n-1)*(sqrt(b2)/h) as written, lerp, then times (1/h)^2.
template <typename Real, typename Real4>
__device__ __forceinline__ Real sph_line_integral(const Real4& sphere, const Real b2, const SmemBlock smem, const int n_table)
{
    const gpu::BoundIter<double> table = smem;
    const Real inv_h = 1.f / sphere.w;
    const Real pos = (n_table - 1) * (sqrt(b2) * inv_h);
    Real w = lerp(pos, table, n_table);
    w *= (inv_h * inv_h);
    return w;
}

} // namespace detail

// ---- no-ops -------------------------------------------------------------------------------

struct Init_null {
    __device__ void operator()(detail::SmemBlock) {}
};

struct RayEntry_null {
    template <typename RayData>
    __device__ void operator()(int, const Ray&, const RayData&, detail::SmemBlock) {}
};
typedef RayEntry_null RayExit_null;

// ---- per-ray state in / out of global arrays ----------------------------------------------

template <typename T>
struct RayEntry_from_array {
    RayEntry_from_array(const T* per_ray_initial) : src_(per_ray_initial) {}

    template <typename RayData>
    __device__ void operator()(int ray, const Ray&, RayData& state, detail::SmemBlock) { state.data = src_[ray]; }

private:
    const T* src_;
};

template <typename T>
struct RayExit_to_array {
    RayExit_to_array(T* per_ray_result) : dst_(per_ray_result) {}

    template <typename RayData>
    __device__ void operator()(int ray, const Ray&, const RayData& state, detail::SmemBlock) { dst_[ray] = state.data; }

private:
    T* dst_;
};

// ---- block set-up --------------------------------------------------------------------------

// Stages `n` values of T from global memory at the start of the user's shared-memory block;
// the traversal kernel synchronises the block after Init returns.
template <typename T>
struct InitGlobalToSmem {
    InitGlobalToSmem(const T* global_values, int n) : src_(global_values), n_(n) {}

    __device__ void operator()(detail::SmemBlock smem)
    {
        gpu::BoundIter<T> dst = smem;
        for (int k = threadIdx.x; k < n_; k += blockDim.x) dst[k] = src_[k];
    }

private:
    const T* src_;
    int n_;
};

// ---- ray/sphere intersection ---------------------------------------------------------------

struct Intersect_sphere_bool {
    template <typename Real4, typename RayData>
    __device__ bool operator()(const Ray& ray, const Real4& sphere, const RayData&, int, detail::SmemBlock)
    {
        typename Real4ToRealMapper<Real4>::type unused_b2, unused_dist;
        return sphere_hit(ray, sphere, unused_b2, unused_dist);
    }
};

// Leaves the squared impact parameter and the distance to the point of closest approach in the
// ray's state for the OnHit functors below.
struct Intersect_sphere_b2dist {
    template <typename Real4, typename RayData>
    __device__ bool operator()(const Ray& ray, const Real4& sphere, RayData& state, int, detail::SmemBlock)
    {
        return sphere_hit(ray, sphere, state.b2, state.dist);
    }
};

// ---- on hit --------------------------------------------------------------------------------

struct OnHit_increment {
    template <typename RayData, typename TPrim>
    __device__ void operator()(int, const Ray&, RayData& state, int, const TPrim&, int, detail::SmemBlock)
    {
        ++state.data;
    }
};

// Column density: state.data += line integral.  Expects the double-precision table at the start
// of the shared-memory block (InitGlobalToSmem<double>).
struct OnHit_sphere_cumulate {
    OnHit_sphere_cumulate(int n_table) : n_table_(n_table) {}

    template <typename RayData, typename Real4>
    __device__ void operator()(int, const Ray&, RayData& state, int, const Real4& sphere, int, detail::SmemBlock smem)
    {
        typedef typename Real4ToRealMapper<Real4>::type Real;
        state.data += detail::sph_line_integral<Real>(sphere, state.b2, smem, n_table_);
    }

private:
    int n_table_;
};

// Hit lists: state.data is the ray's write cursor (initialised to its offset by
// RayEntry_from_array), one (index, integral, distance) record per hit.
template <typename IndexType, typename Real>
struct OnHit_sphere_individual {
    OnHit_sphere_individual(IndexType* hit_indices, Real* hit_integrals, Real* hit_distances, int n_table)
        : idx_(hit_indices), w_(hit_integrals), dist_(hit_distances), n_table_(n_table) {}

    template <typename RayData, typename Real4>
    __device__ void operator()(int, const Ray&, RayData& state, int prim, const Real4& sphere, int, detail::SmemBlock smem)
    {
        const Real w = detail::sph_line_integral<Real>(sphere, state.b2, smem, n_table_);
        const int slot = state.data++;
        idx_[slot] = prim;
        w_[slot] = w;
        dist_[slot] = state.dist;
    }

private:
    IndexType* idx_;
    Real* w_;
    Real* dist_;
    int n_table_;
};

} // namespace grace
