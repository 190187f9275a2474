// grace/generic/functors/aabb.h -- bounding box of an SPH particle: centre -/+ smoothing length
// per axis (reference behaviour: generic/functors/aabb.h:9-43).
#pragma once
#include "grace/types.h"

namespace grace {

namespace detail {

// Midpoint formed in double so that two large floats cannot overflow or lose the last bit.
GRACE_HOST_DEVICE float midpoint_f(const float lo, const float hi)
{
    return (static_cast<double>(lo) + hi) / 2.;
}

GRACE_HOST_DEVICE float3 AABB_centroid(const float3 bot, const float3 top)
{
    float3 mid;
    mid.x = midpoint_f(bot.x, top.x);
    mid.y = midpoint_f(bot.y, top.y);
    mid.z = midpoint_f(bot.z, top.z);
    return mid;
}

} // namespace detail

// Works for float4 and double4: the subtraction/addition is made in the sphere's own precision
// and narrowed on assignment to the float3 corners.
struct AABBSphere {
    template <typename Real4>
    GRACE_HOST_DEVICE void operator()(Real4 s, float3* bot, float3* top) const
    {
        bot->x = s.x - s.w;
        bot->y = s.y - s.w;
        bot->z = s.z - s.w;
        top->x = s.x + s.w;
        top->y = s.y + s.w;
        top->z = s.z + s.w;
    }
};

} // namespace grace
