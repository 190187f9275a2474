// grace/generic/functors/centroid.h -- centroid functors (reference: generic/functors/centroid.h:17-40).
#pragma once
#include "grace/generic/functors/aabb.h"

namespace grace {

template <typename TPrimitive, typename AABBFunc>
struct PrimitiveCentroid {
    GRACE_HOST_DEVICE float3 operator()(TPrimitive primitive) const
    {
        float3 bot, top;
        AABBFunc()(primitive, &bot, &top);
        return detail::AABB_centroid(bot, top);
    }
};

struct CentroidSphere {
    template <typename Real4>
    GRACE_HOST_DEVICE float3 operator()(Real4 sphere) const
    {
        float3 c; c.x = sphere.x; c.y = sphere.y; c.z = sphere.z;
        return c;
    }
};

} // namespace grace
