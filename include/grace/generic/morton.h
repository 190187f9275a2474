// grace/generic/morton.h -- host/device Morton keys (reference: generic/morton.h:14-55).
#pragma once
#include "grace/generic/bits.h"

namespace grace {

GRACE_HOST_DEVICE uinteger32 morton_key(const uinteger32 x, const uinteger32 y, const uinteger32 z)
{
    return detail::space_by_two_10bit(z) << 2 | detail::space_by_two_10bit(y) << 1 | detail::space_by_two_10bit(x);
}

GRACE_HOST_DEVICE uinteger64 morton_key(const uinteger64 x, const uinteger64 y, const uinteger64 z)
{
    return detail::space_by_two_21bit(z) << 2 | detail::space_by_two_21bit(y) << 1 | detail::space_by_two_21bit(x);
}

// 30-bit key of a point in (0, 1)^3.
GRACE_HOST_DEVICE uinteger32 morton_key(const float x, const float y, const float z)
{
    const unsigned int span = (1u << 10) - 1;
    return morton_key(static_cast<uinteger32>(span * x), static_cast<uinteger32>(span * y),
                      static_cast<uinteger32>(span * z));
}

// 63-bit key of a point in (0, 1)^3.
GRACE_HOST_DEVICE uinteger64 morton_key(const double x, const double y, const double z)
{
    const unsigned int span = (1u << 21) - 1;
    return morton_key(static_cast<uinteger64>(span * x), static_cast<uinteger64>(span * y),
                      static_cast<uinteger64>(span * z));
}

} // namespace grace
