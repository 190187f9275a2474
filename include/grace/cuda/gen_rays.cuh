// grace/cuda/gen_rays.cuh -- ray generators (reference: cuda/gen_rays.cuh:26-399).
// Each has a raw-pointer and a container form; the container form grows d_rays if needed.
#pragma once
#include "grace/cuda/sort.cuh"      // as in the reference (kernels/gen_rays.cuh:8): sort_by_distance comes along
#include "grace/cuda/util/extrema.cuh"
#include "grace/device_vector.h"
#include "grace/ray.h"

namespace grace {
namespace detail {
inline grace_b200_ray* rp(Ray* r) { return reinterpret_cast<grace_b200_ray*>(r); }
template <typename V3> inline void to3(const V3& v, float out[3]) { out[0] = (float)v.x; out[1] = (float)v.y; out[2] = (float)v.z; }
template <typename P> struct point_stride { enum { value = sizeof(P) / sizeof(float) }; };
}

template <typename Real>
GRACE_HOST void uniform_random_rays(Ray* const d_rays_ptr, const size_t N_rays, const Real ox, const Real oy,
                                    const Real oz, const Real length, const unsigned long long seed = 1234)
{
    GRACE_B200_CHECK(grace_b200_uniform_random_rays(detail::context(), detail::rp(d_rays_ptr), N_rays, (float)ox,
                                                    (float)oy, (float)oz, (float)length, -1, seed, nullptr));
}
template <typename RayVec, typename Real>
GRACE_HOST void uniform_random_rays(RayVec& d_rays, const Real ox, const Real oy, const Real oz, const Real length,
                                    const unsigned long long seed = 1234)
{
    uniform_random_rays(detail::raw(d_rays.data()), d_rays.size(), ox, oy, oz, length, seed);
}

template <typename Real>
GRACE_HOST void uniform_random_rays_single_octant(Ray* const d_rays_ptr, const size_t N_rays, const Real ox,
                                                  const Real oy, const Real oz, const Real length,
                                                  const enum Octants octant = PPP, const unsigned long long seed = 1234)
{
    GRACE_B200_CHECK(grace_b200_uniform_random_rays(detail::context(), detail::rp(d_rays_ptr), N_rays, (float)ox,
                                                    (float)oy, (float)oz, (float)length, (int)octant, seed, nullptr));
}
template <typename RayVec, typename Real>
GRACE_HOST void uniform_random_rays_single_octant(RayVec& d_rays, const Real ox, const Real oy, const Real oz,
                                                  const Real length, const enum Octants octant = PPP,
                                                  const unsigned long long seed = 1234)
{
    uniform_random_rays_single_octant(detail::raw(d_rays.data()), d_rays.size(), ox, oy, oz, length, octant, seed);
}

// Rays from (ox, oy, oz) to each point.  Throws std::invalid_argument for an unknown sort type.
template <typename Real, typename PointType>
GRACE_HOST void one_to_many_rays(Ray* const d_rays_ptr, const size_t N_rays, const Real ox, const Real oy,
                                 const Real oz, const PointType* const d_points_ptr,
                                 const enum RaySortType sort_type = DirectionSort)
{
    GRACE_B200_CHECK(grace_b200_one_to_many_rays(detail::context(), detail::rp(d_rays_ptr), N_rays, (float)ox, (float)oy,
                                                 (float)oz, reinterpret_cast<const float*>(d_points_ptr),
                                                 detail::point_stride<PointType>::value, (int)sort_type, nullptr,
                                                 nullptr, nullptr));
}
template <typename RayVec, typename Real, typename PointVec>
GRACE_HOST void one_to_many_rays(RayVec& d_rays, const Real ox, const Real oy, const Real oz, const PointVec& d_points,
                                 const enum RaySortType sort_type = DirectionSort)
{
    if (d_rays.size() < d_points.size()) d_rays.resize(d_points.size());
    one_to_many_rays(detail::raw(d_rays.data()), d_points.size(), ox, oy, oz, detail::raw(d_points.data()), sort_type);
}
// End-point sort with explicit bounds for the end points.
template <typename Real, typename Real3, typename PointType>
GRACE_HOST void one_to_many_rays(Ray* const d_rays_ptr, const size_t N_rays, const Real ox, const Real oy,
                                 const Real oz, const PointType* const d_points_ptr, const Real3 AABB_bot,
                                 const Real3 AABB_top)
{
    float b[3], t[3];
    detail::to3(AABB_bot, b); detail::to3(AABB_top, t);
    GRACE_B200_CHECK(grace_b200_one_to_many_rays(detail::context(), detail::rp(d_rays_ptr), N_rays, (float)ox, (float)oy,
                                                 (float)oz, reinterpret_cast<const float*>(d_points_ptr),
                                                 detail::point_stride<PointType>::value, GRACE_B200_ENDPOINT_SORT, b, t,
                                                 nullptr));
}
template <typename RayVec, typename Real, typename Real3, typename PointVec>
GRACE_HOST void one_to_many_rays(RayVec& d_rays, const Real ox, const Real oy, const Real oz, const PointVec& d_points,
                                 const Real3 AABB_bot, const Real3 AABB_top)
{
    if (d_rays.size() < d_points.size()) d_rays.resize(d_points.size());
    one_to_many_rays(detail::raw(d_rays.data()), d_points.size(), ox, oy, oz, detail::raw(d_points.data()), AABB_bot,
                     AABB_top);
}

template <typename Real, typename Real3>
GRACE_HOST void plane_parallel_random_rays(Ray* const d_rays_ptr, const int width, const int height, const Real3 base,
                                           const Real3 w, const Real3 h, const Real length,
                                           const unsigned long long seed = 1234)
{
    float b[3], ww[3], hh[3];
    detail::to3(base, b); detail::to3(w, ww); detail::to3(h, hh);
    GRACE_B200_CHECK(grace_b200_plane_parallel_random_rays(detail::context(), detail::rp(d_rays_ptr), width, height, b,
                                                           ww, hh, (float)length, seed, nullptr));
}
template <typename RayVec, typename Real, typename Real3>
GRACE_HOST void plane_parallel_random_rays(RayVec& d_rays, const int width, const int height, const Real3 base,
                                           const Real3 w, const Real3 h, const Real length,
                                           const unsigned long long seed = 1234)
{
    const size_t n = (size_t)width * height;
    if (d_rays.size() < n) d_rays.resize(n);
    plane_parallel_random_rays(detail::raw(d_rays.data()), width, height, base, w, h, length, seed);
}

template <typename Real, typename Real3>
GRACE_HOST void orthographic_projection_rays(Ray* const d_rays_ptr, const int resolution_x, const int resolution_y,
                                             const Real3 camera_position, const Real3 look_at, const Real3 view_up,
                                             const Real vertical_extent, const Real length)
{
    float c[3], l[3], u[3];
    detail::to3(camera_position, c); detail::to3(look_at, l); detail::to3(view_up, u);
    GRACE_B200_CHECK(grace_b200_orthographic_projection_rays(detail::context(), detail::rp(d_rays_ptr), resolution_x,
                                                             resolution_y, c, l, u, (float)vertical_extent,
                                                             (float)length, nullptr));
}
template <typename RayVec, typename Real, typename Real3>
GRACE_HOST void orthographic_projection_rays(RayVec& d_rays, const int resolution_x, const int resolution_y,
                                             const Real3 camera_position, const Real3 look_at, const Real3 view_up,
                                             const Real vertical_extent, const Real length)
{
    const size_t n = (size_t)resolution_x * resolution_y;
    if (d_rays.size() < n) d_rays.resize(n);
    orthographic_projection_rays(detail::raw(d_rays.data()), resolution_x, resolution_y, camera_position, look_at,
                                 view_up, vertical_extent, length);
}

template <typename Real, typename Real3>
GRACE_HOST void pinhole_camera_rays(Ray* const d_rays_ptr, const int resolution_x, const int resolution_y,
                                    const Real3 camera_position, const Real3 look_at, const Real3 view_up,
                                    const Real FOVy, const Real length)
{
    float c[3], l[3], u[3];
    detail::to3(camera_position, c); detail::to3(look_at, l); detail::to3(view_up, u);
    GRACE_B200_CHECK(grace_b200_pinhole_camera_rays(detail::context(), detail::rp(d_rays_ptr), resolution_x,
                                                    resolution_y, c, l, u, (float)FOVy, (float)length, nullptr));
}
template <typename RayVec, typename Real, typename Real3>
GRACE_HOST void pinhole_camera_rays(RayVec& d_rays, const int resolution_x, const int resolution_y,
                                    const Real3 camera_position, const Real3 look_at, const Real3 view_up,
                                    const Real FOVy, const Real length)
{
    const size_t n = (size_t)resolution_x * resolution_y;
    if (d_rays.size() < n) d_rays.resize(n);
    pinhole_camera_rays(detail::raw(d_rays.data()), resolution_x, resolution_y, camera_position, look_at, view_up,
                        FOVy, length);
}

// HEALPix NESTED pixel-centre rays [first_pixel, first_pixel + n) at resolution nside
// (the direction set of RayVectorGeneration/src/chealpix/chealpix.c:459-467).
template <typename RayVec, typename Real>
GRACE_HOST void healpix_rays(RayVec& d_rays, const long nside, const long first_pixel, const size_t n, const Real ox,
                             const Real oy, const Real oz, const Real length)
{
    if (d_rays.size() < n) d_rays.resize(n);
    GRACE_B200_CHECK(grace_b200_healpix_rays(detail::context(), detail::rp(detail::raw(d_rays.data())), n, nside,
                                             first_pixel, (float)ox, (float)oy, (float)oz, (float)length, nullptr));
}

} // namespace grace
