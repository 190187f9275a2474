// grace/generic/intersect.h -- ray/sphere test usable on the host (reference:
// generic/intersect.h:10-55).  A ray hits when its closest approach lies inside the sphere
// and within [0, length) along the ray; origins/termini inside the sphere beyond closest
// approach count as misses.
#pragma once
#include "grace/ray.h"
#include "grace/types.h"

namespace grace {

template <typename Real4, typename Real>
GRACE_HOST_DEVICE bool sphere_hit(const Ray& ray, const Real4& sphere, Real& b2, Real& dot_p)
{
    const Real px = sphere.x - ray.ox, py = sphere.y - ray.oy, pz = sphere.z - ray.oz;
    const Real rx = ray.dx, ry = ray.dy, rz = ray.dz;      // already normalised
    dot_p = px * rx + py * ry + pz * rz;                   // distance to closest approach
    const Real bx = px - dot_p * rx, by = py - dot_p * ry, bz = pz - dot_p * rz;
    b2 = bx * bx + by * by + bz * bz;                      // impact parameter squared
    if (b2 >= sphere.w * sphere.w) return false;
    if (dot_p < 0.0f) return false;
    if (dot_p >= ray.length) return false;
    return true;
}

} // namespace grace
