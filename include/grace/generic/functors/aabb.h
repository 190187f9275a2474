// grace/generic/functors/aabb.h -- AABB of an SPH sphere (reference: generic/functors/aabb.h:9-43).
#pragma once
#include "grace/types.h"

namespace grace {

struct AABBSphere {
    template <typename Real4>
    GRACE_HOST_DEVICE void operator()(Real4 sphere, float3* bot, float3* top) const
    {
        bot->x = sphere.x - sphere.w; top->x = sphere.x + sphere.w;
        bot->y = sphere.y - sphere.w; top->y = sphere.y + sphere.w;
        bot->z = sphere.z - sphere.w; top->z = sphere.z + sphere.w;
    }
};

namespace detail {
GRACE_HOST_DEVICE float3 AABB_centroid(const float3 bot, const float3 top)
{
    float3 c;
    c.x = (static_cast<double>(bot.x) + top.x) / 2.;
    c.y = (static_cast<double>(bot.y) + top.y) / 2.;
    c.z = (static_cast<double>(bot.z) + top.z) / 2.;
    return c;
}
} // namespace detail

} // namespace grace
