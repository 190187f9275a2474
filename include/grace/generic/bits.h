// grace/generic/bits.h -- bit spreading for Morton keys (reference: generic/bits.h:12-46).
#pragma once
#include "grace/types.h"

namespace grace {

template <typename T>
GRACE_HOST_DEVICE int sgn(T val) { return (T(0) < val) - (val < T(0)); }

namespace detail {

// Insert two zero bits after each of the low 10 bits of x.
template <typename UInteger>
GRACE_HOST_DEVICE uinteger32 space_by_two_10bit(const UInteger x)
{
    uinteger32 v = static_cast<uinteger32>(x) & 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// Insert two zero bits after each of the low 21 bits of x (63 bits in total).
template <typename UInteger>
GRACE_HOST_DEVICE uinteger64 space_by_two_21bit(const UInteger x)
{
    uinteger64 v = static_cast<uinteger64>(x) & 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

} // namespace detail
} // namespace grace
