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

class Init_null {
public:
    __device__ void operator()(const gpu::BoundIter<char>) {}
};

class RayEntry_null {
public:
    template <typename RayData>
    __device__ void operator()(const int, const Ray&, const RayData&, const gpu::BoundIter<char>) {}
};
typedef RayEntry_null RayExit_null;

template <typename T>
class RayEntry_from_array {
    const T* const inits;
public:
    RayEntry_from_array(const T* const ray_data_inits) : inits(ray_data_inits) {}
    template <typename RayData>
    __device__ void operator()(const int ray_idx, const Ray&, RayData& ray_data, const gpu::BoundIter<char>)
    {
        ray_data.data = inits[ray_idx];
    }
};

template <typename T>
class RayExit_to_array {
    T* const store;
public:
    RayExit_to_array(T* const ray_data_store) : store(ray_data_store) {}
    template <typename RayData>
    __device__ void operator()(const int ray_idx, const Ray&, const RayData& ray_data, const gpu::BoundIter<char>)
    {
        store[ray_idx] = ray_data.data;
    }
};

// Copies `count` values from global to the user's shared-memory block (the kernel synchronises after).
template <typename T>
class InitGlobalToSmem {
    const T* const data_global;
    const int count;
public:
    InitGlobalToSmem(const T* const global_addr, const int count) : data_global(global_addr), count(count) {}
    __device__ void operator()(const gpu::BoundIter<char> smem_iter)
    {
        gpu::BoundIter<T> T_iter = smem_iter;
        for (int i = threadIdx.x; i < count; i += blockDim.x) T_iter[i] = data_global[i];
    }
};

class Intersect_sphere_bool {
public:
    template <typename Real4, typename RayData>
    __device__ bool operator()(const Ray& ray, const Real4& sphere, const RayData&, const int, const gpu::BoundIter<char>)
    {
        typedef typename Real4ToRealMapper<Real4>::type Real;
        Real b2, dist;
        return sphere_hit(ray, sphere, b2, dist);
    }
};

class Intersect_sphere_b2dist {
public:
    template <typename Real4, typename RayData>
    __device__ bool operator()(const Ray& ray, const Real4& sphere, RayData& ray_data, const int, const gpu::BoundIter<char>)
    {
        return sphere_hit(ray, sphere, ray_data.b2, ray_data.dist);
    }
};

class OnHit_increment {
public:
    template <typename RayData, typename TPrim>
    __device__ void operator()(const int, const Ray&, RayData& ray_data, const int, const TPrim&, const int,
                               const gpu::BoundIter<char>)
    {
        ++ray_data.data;
    }
};

// Accumulates kernel line integrals; the double-precision table sits at the start of the user's
// shared-memory block (InitGlobalToSmem<double>).
class OnHit_sphere_cumulate {
    const int N_table;
public:
    OnHit_sphere_cumulate(const int N_table) : N_table(N_table) {}
    template <typename RayData, typename Real4>
    __device__ void operator()(const int, const Ray&, RayData& ray_data, const int, const Real4& sphere, const int,
                               const gpu::BoundIter<char> smem_iter)
    {
        typedef typename Real4ToRealMapper<Real4>::type Real;
        gpu::BoundIter<double> Wk_lookup = smem_iter;
        Real ir = 1.f / sphere.w;
        Real b = (N_table - 1) * (sqrt(ray_data.b2) * ir);
        Real integral = lerp(b, Wk_lookup, N_table);
        integral *= (ir * ir);
        ray_data.data += integral;
    }
};

template <typename IndexType, typename Real>
class OnHit_sphere_individual {
    IndexType* const indices;
    Real* const integrals;
    Real* const distances;
    const int N_table;
public:
    OnHit_sphere_individual(IndexType* const indices, Real* const integrals, Real* const distances, const int N_table)
        : indices(indices), integrals(integrals), distances(distances), N_table(N_table) {}
    template <typename RayData, typename Real4>
    __device__ void operator()(const int, const Ray&, RayData& ray_data, const int sphere_idx, const Real4& sphere,
                               const int, const gpu::BoundIter<char> smem_iter)
    {
        gpu::BoundIter<double> Wk_lookup = smem_iter;
        Real ir = 1.f / sphere.w;
        Real b = (N_table - 1) * (sqrt(ray_data.b2) * ir);
        Real integral = lerp(b, Wk_lookup, N_table);
        integral *= (ir * ir);
        indices[ray_data.data] = sphere_idx;
        integrals[ray_data.data] = integral;
        distances[ray_data.data] = ray_data.dist;
        ++ray_data.data;
    }
};

} // namespace grace
