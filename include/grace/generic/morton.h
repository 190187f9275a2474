// grace/generic/morton.h -- host/device Morton keys: x in the lowest bit of each triple, then y,
// then z (reference behaviour: generic/morton.h:14-55).
#pragma once
#include "grace/generic/bits.h"

namespace grace {

namespace detail {

// Grid cell of a coordinate in (0, 1): (2^bits - 1) * u, truncated.  The product is formed in
// the coordinate's own precision (float for 30-bit keys, double for 63-bit keys).
template <typename Key, int BITS, typename Real>
GRACE_HOST_DEVICE Key morton_cell(const Real u)
{
    const unsigned int cells_minus_one = (1u << BITS) - 1;
    return static_cast<Key>(cells_minus_one * u);
}

} // namespace detail

// Integer cells in, key out: 10 bits per axis ...
GRACE_HOST_DEVICE uinteger32 morton_key(const uinteger32 x, const uinteger32 y, const uinteger32 z)
{
    const uinteger32 sx = detail::space_by_two_10bit(x);
    const uinteger32 sy = detail::space_by_two_10bit(y);
    const uinteger32 sz = detail::space_by_two_10bit(z);
    return sx | (sy << 1) | (sz << 2);
}

// ... or 21 bits per axis.
GRACE_HOST_DEVICE uinteger64 morton_key(const uinteger64 x, const uinteger64 y, const uinteger64 z)
{
    const uinteger64 sx = detail::space_by_two_21bit(x);
    const uinteger64 sy = detail::space_by_two_21bit(y);
    const uinteger64 sz = detail::space_by_two_21bit(z);
    return sx | (sy << 1) | (sz << 2);
}

// Points of the unit cube: float coordinates give the 30-bit key, double coordinates the 63-bit key.
GRACE_HOST_DEVICE uinteger32 morton_key(const float x, const float y, const float z)
{
    return morton_key(detail::morton_cell<uinteger32, 10>(x), detail::morton_cell<uinteger32, 10>(y),
                      detail::morton_cell<uinteger32, 10>(z));
}

GRACE_HOST_DEVICE uinteger64 morton_key(const double x, const double y, const double z)
{
    return morton_key(detail::morton_cell<uinteger64, 21>(x), detail::morton_cell<uinteger64, 21>(y),
                      detail::morton_cell<uinteger64, 21>(z));
}

} // namespace grace
