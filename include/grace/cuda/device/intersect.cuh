// grace/cuda/device/intersect.cuh -- the two-box slab test of the generic traversal (reference:
// cuda/device/intersect.cuh:10-40 + cuda/device/intrinsics.cuh:8-52), in the arithmetic nvcc gives
// the reference: t = (plane - o) * (1/d) (subtract, then multiply), per-axis fmin/fmax, the
// three-input stages as SIGNED INTEGER min/max on the float bit patterns, tmax >= tmin.
#pragma once
#include "grace/types.h"

namespace grace {
namespace detail {

__device__ __forceinline__ bool slab_hit(float bx, float tx, float by, float ty, float bz, float tz,
                                         float ox, float oy, float oz, float ix, float iy, float iz, float len)
{
    const float tbx = __fmul_rn(__fsub_rn(bx, ox), ix), ttx = __fmul_rn(__fsub_rn(tx, ox), ix);
    const float tby = __fmul_rn(__fsub_rn(by, oy), iy), tty = __fmul_rn(__fsub_rn(ty, oy), iy);
    const float tbz = __fmul_rn(__fsub_rn(bz, oz), iz), ttz = __fmul_rn(__fsub_rn(tz, oz), iz);
    const int zmin = max(min(__float_as_int(tbz), __float_as_int(ttz)), 0);
    const int zmax = min(max(__float_as_int(tbz), __float_as_int(ttz)), __float_as_int(len));
    const int tmin = max(max(__float_as_int(fminf(tbx, ttx)), __float_as_int(fminf(tby, tty))), zmin);
    const int tmax = min(min(__float_as_int(fmaxf(tbx, ttx)), __float_as_int(fmaxf(tby, tty))), zmax);
    return __int_as_float(tmax) >= __int_as_float(tmin);
}

} // namespace detail

// hit right + 2 * hit left (device/intersect.cuh:16-39)
__device__ __forceinline__ int AABBs_hit(const float3 invd, const float3 origin, const float length,
                                         const float4 AABB_L, const float4 AABB_R, const float4 AABB_LR)
{
    const bool l = detail::slab_hit(AABB_L.x, AABB_L.y, AABB_L.z, AABB_L.w, AABB_LR.x, AABB_LR.y,
                                    origin.x, origin.y, origin.z, invd.x, invd.y, invd.z, length);
    const bool r = detail::slab_hit(AABB_R.x, AABB_R.y, AABB_R.z, AABB_R.w, AABB_LR.z, AABB_LR.w,
                                    origin.x, origin.y, origin.z, invd.x, invd.y, invd.z, length);
    return (int)r + 2 * (int)l;
}

} // namespace grace
