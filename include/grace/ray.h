// grace/ray.h -- reference: include/grace/ray.h:5-10.  Layout is part of the ABI
// (grace_b200_ray in grace_b200.h is the same 7 floats).
#pragma once
namespace grace {
struct Ray { float dx, dy, dz, ox, oy, oz, length; };
}
