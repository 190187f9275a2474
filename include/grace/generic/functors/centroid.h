// grace/generic/functors/centroid.h -- what the Morton-key stage asks of a primitive: one
// float3 "where is it" (reference behaviour: generic/functors/centroid.h:17-40).
#pragma once
#include "grace/generic/functors/aabb.h"

namespace grace {

// Centre of whatever box AABBFunc reports for the primitive.  AABBFunc must be default
// constructible; it is called as AABBFunc()(prim, &lo, &hi).
template <typename TPrimitive, typename AABBFunc>
struct PrimitiveCentroid {
    GRACE_HOST_DEVICE float3 operator()(TPrimitive prim) const
    {
        float3 lo, hi;
        AABBFunc box_of;
        box_of(prim, &lo, &hi);
        return detail::AABB_centroid(lo, hi);
    }
};

// A sphere {x, y, z, radius} sits at its xyz; double4 inputs are narrowed to float here, which
// is also where the reference narrows them.
struct CentroidSphere {
    template <typename Real4>
    GRACE_HOST_DEVICE float3 operator()(Real4 s) const
    {
        return make_float3((float)s.x, (float)s.y, (float)s.z);
    }
};

} // namespace grace
