// grace/generic/functors/albvh.h -- host-callable delta functors (reference:
// generic/functors/albvh.h:17-126).  delta(i) relates elements i and i+1; indices outside
// [0, n-1) return the maximum value (UINT_MAX / +inf).  The device build path evaluates the
// same definitions inside grace_b200_deltas_* (csrc/albvh.cu).
#pragma once
#include <cstddef>
#include <iterator>
#include <limits>
#include "grace/types.h"

namespace grace {

namespace detail {
GRACE_HOST_DEVICE float delta_infinity()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7f800000);
#else
    return std::numeric_limits<float>::infinity();
#endif
}
} // namespace detail

struct DeltaXOR {
    GRACE_HOST_DEVICE uinteger32 operator()(const int i, const uinteger32* keys, const size_t n) const
    {
        if (i < 0 || (size_t)i + 1 >= n) return uinteger32(-1);
        return keys[i] ^ keys[i + 1];
    }
    GRACE_HOST_DEVICE uinteger64 operator()(const int i, const uinteger64* keys, const size_t n) const
    {
        if (i < 0 || (size_t)i + 1 >= n) return uinteger64(-1);
        return keys[i] ^ keys[i + 1];
    }
};

template <typename PrimitiveIter, typename CentroidFunc>
struct DeltaEuclidean {
    GRACE_HOST_DEVICE float operator()(const int i, PrimitiveIter prims, const size_t n) const
    {
        if (i < 0 || (size_t)i + 1 >= n) return detail::delta_infinity();
        typename std::iterator_traits<PrimitiveIter>::value_type a = prims[i], b = prims[i + 1];
        return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z);
    }
};

template <typename PrimitiveIter, typename AABBFunc>
struct DeltaSurfaceArea {
    GRACE_HOST_DEVICE float operator()(const int i, PrimitiveIter prims, const size_t n) const
    {
        if (i < 0 || (size_t)i + 1 >= n) return detail::delta_infinity();
        float3 bi, ti, bj, tj;
        AABBFunc()(prims[i], &bi, &ti);
        AABBFunc()(prims[i + 1], &bj, &tj);
        const float Lx = (ti.x > tj.x ? ti.x : tj.x) - (bi.x < bj.x ? bi.x : bj.x);
        const float Ly = (ti.y > tj.y ? ti.y : tj.y) - (bi.y < bj.y ? bi.y : bj.y);
        const float Lz = (ti.z > tj.z ? ti.z : tj.z) - (bi.z < bj.z ? bi.z : bj.z);
        return (Lx * Ly) + (Lx * Lz) + (Ly * Lz);
    }
};

} // namespace grace
