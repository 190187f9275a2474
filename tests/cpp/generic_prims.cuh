// generic_prims.cuh -- a user-defined primitive (triangle) with its functors, written ONLY against
// the API GRACE and this repo share (grace::Ray, grace::gpu::BoundIter, the functor signatures of
// cuda/functors/trace.cuh).  Included by tests/cpp/generic_test.cu (this repo's headers) and by
// oracle/ref_generic_driver.cu (the reference's headers): same source, two implementations.
// Modelled on the reference's triangle demo (tests/profile_trace_triangle/tris_trace.cuh:11-73,
// tris_tree.cuh:17-30); test infrastructure, not product code.
#pragma once
#include "grace/ray.h"
#include "grace/cuda/util/bound_iter.cuh"

#include <vector>

struct Tri { float3 v, e1, e2; };     // a vertex and two edges

static inline unsigned gp_hash(unsigned a)
{
    a = (a + 0x7ed55d16) + (a << 12); a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);  a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);  a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
static inline float gp_u01(unsigned i, unsigned k) { return (gp_hash(i * 16u + k) >> 8) * (1.0f / 16777216.0f); }

static inline std::vector<Tri> make_tris(size_t n, float size)
{
    std::vector<Tri> t(n);
    for (size_t i = 0; i < n; ++i) {
        const unsigned u = (unsigned)i;
        t[i].v = make_float3(gp_u01(u, 0), gp_u01(u, 1), gp_u01(u, 2));
        t[i].e1 = make_float3(size * (gp_u01(u, 3) - 0.5f), size * (gp_u01(u, 4) - 0.5f), size * (gp_u01(u, 5) - 0.5f));
        t[i].e2 = make_float3(size * (gp_u01(u, 6) - 0.5f), size * (gp_u01(u, 7) - 0.5f), size * (gp_u01(u, 8) - 0.5f));
    }
    return t;
}

// rays from a point towards a jittered grid of directions (unit length directions)
static inline std::vector<grace::Ray> make_rays(size_t n)
{
    std::vector<grace::Ray> r(n);
    for (size_t i = 0; i < n; ++i) {
        const unsigned u = (unsigned)i + 77777u;
        float dx = 2.f * gp_u01(u, 0) - 1.f, dy = 2.f * gp_u01(u, 1) - 1.f, dz = 2.f * gp_u01(u, 2) - 1.f;
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz + 1e-12f);
        r[i].dx = dx * inv; r[i].dy = dy * inv; r[i].dz = dz * inv;
        r[i].ox = 0.5f; r[i].oy = 0.5f; r[i].oz = 0.5f;
        r[i].length = 2.0f;
    }
    return r;
}

struct TriAABB {
    __host__ __device__ void operator()(const Tri& t, float3* bot, float3* top) const
    {
        const float3 a = t.v, b = make_float3(t.v.x + t.e1.x, t.v.y + t.e1.y, t.v.z + t.e1.z),
                     c = make_float3(t.v.x + t.e2.x, t.v.y + t.e2.y, t.v.z + t.e2.z);
        bot->x = fminf(a.x, fminf(b.x, c.x)); top->x = fmaxf(a.x, fmaxf(b.x, c.x));
        bot->y = fminf(a.y, fminf(b.y, c.y)); top->y = fmaxf(a.y, fmaxf(b.y, c.y));
        bot->z = fminf(a.z, fminf(b.z, c.z)); top->z = fmaxf(a.z, fmaxf(b.z, c.z));
    }
};

struct TriCentroid {
    __host__ __device__ float3 operator()(const Tri& t) const
    {
        float3 bot, top;
        TriAABB()(t, &bot, &top);
        return make_float3(0.5f * (bot.x + top.x), 0.5f * (bot.y + top.y), 0.5f * (bot.z + top.z));
    }
};

// Moeller-Trumbore, written with explicit fmaf so host and device agree bit for bit
__host__ __device__ inline bool tri_intersect(const grace::Ray& ray, const Tri& t, float* t_out)
{
    const float px = fmaf(ray.dy, t.e2.z, -ray.dz * t.e2.y), py = fmaf(ray.dz, t.e2.x, -ray.dx * t.e2.z),
                pz = fmaf(ray.dx, t.e2.y, -ray.dy * t.e2.x);
    const float det = fmaf(t.e1.x, px, fmaf(t.e1.y, py, t.e1.z * pz));
    if (fabsf(det) < 1e-12f) return false;
    const float inv = 1.0f / det;
    const float sx = ray.ox - t.v.x, sy = ray.oy - t.v.y, sz = ray.oz - t.v.z;
    const float u = fmaf(sx, px, fmaf(sy, py, sz * pz)) * inv;
    if (u < 0.f || u > 1.f) return false;
    const float qx = fmaf(sy, t.e1.z, -sz * t.e1.y), qy = fmaf(sz, t.e1.x, -sx * t.e1.z), qz = fmaf(sx, t.e1.y, -sy * t.e1.x);
    const float v = fmaf(ray.dx, qx, fmaf(ray.dy, qy, ray.dz * qz)) * inv;
    if (v < 0.f || u + v > 1.f) return false;
    *t_out = fmaf(t.e2.x, qx, fmaf(t.e2.y, qy, t.e2.z * qz)) * inv;
    return true;
}

struct RayData_tri { int data; float t_min; };      // .data = closest triangle so far

struct RayIntersect_tri {
    __device__ bool operator()(const grace::Ray& ray, const Tri& tri, RayData_tri& rd, const int,
                               const grace::gpu::BoundIter<char>) const
    {
        float t;
        if (tri_intersect(ray, tri, &t) && t <= rd.t_min && t >= 1e-6f) { rd.t_min = t; return true; }
        return false;
    }
};
struct OnHit_tri {
    __device__ void operator()(const int, const grace::Ray&, RayData_tri& rd, const int tri_idx, const Tri&, const int,
                               const grace::gpu::BoundIter<char>) const
    {
        rd.data = tri_idx;
    }
};
struct RayEntry_tri {
    __device__ void operator()(const int, const grace::Ray& ray, RayData_tri& rd, const grace::gpu::BoundIter<char>) const
    {
        rd.data = -1;
        rd.t_min = ray.length;
    }
};
// counts every triangle the ray crosses within its length (order-independent: checks coverage)
struct RayData_cnt { int data; };
struct RayIntersect_any {
    __device__ bool operator()(const grace::Ray& ray, const Tri& tri, const RayData_cnt&, const int,
                               const grace::gpu::BoundIter<char>) const
    {
        float t;
        return tri_intersect(ray, tri, &t) && t >= 0.f && t < ray.length;
    }
};
