// grace/generic/bits.h -- bit spreading for Morton keys.  Same results as the reference's
// space_by_two_10bit / space_by_two_21bit (generic/bits.h:24-46); here both are instances of one
// template whose masks are derived at compile time instead of being spelt out.
#pragma once
#include "grace/types.h"

namespace grace {

template <typename T>
GRACE_HOST_DEVICE int sgn(T val) { return (T(0) < val) - (val < T(0)); }

namespace detail {

// The mask that keeps `n_ones` bits laid out as runs of `run` ones, one run every 3*run bits,
// starting at bit 0.  run = 1 is the final "every third bit" pattern (0x09249249 for 10 bits).
template <typename U>
constexpr U spread3_mask(int run, int n_ones)
{
    U m = 0;
    for (int placed = 0, bit = 0; placed < n_ones; bit += 3 * run)
        for (int k = 0; k < run && placed < n_ones; ++k, ++placed) m |= U(1) << (bit + k);
    return m;
}

// Moves bit b of the low BITS bits of v to bit 3*b by doubling the gaps: at each level a run of
// 2*RUN bits is split into two runs of RUN bits, the upper one moved up by 2*RUN.  The recursion
// starts from the narrowest level (RUN = 1) and bottoms out once a run would hold all BITS bits.
template <typename U, int BITS, int RUN, bool WHOLE = (RUN >= BITS)>
struct Spread3 {
    GRACE_HOST_DEVICE static U apply(U v)
    {
        v = Spread3<U, BITS, 2 * RUN>::apply(v);
        return (v | (v << (2 * RUN))) & spread3_mask<U>(RUN, BITS);
    }
};
template <typename U, int BITS, int RUN>
struct Spread3<U, BITS, RUN, true> {
    GRACE_HOST_DEVICE static U apply(U v) { return v & ((U(1) << BITS) - 1); }
};

// x's low 10 bits, two zero bits after each (30 bits).
template <typename UInteger>
GRACE_HOST_DEVICE uinteger32 space_by_two_10bit(const UInteger x)
{
    return Spread3<uinteger32, 10, 1>::apply(static_cast<uinteger32>(x));
}

// x's low 21 bits, two zero bits after each (63 bits).
template <typename UInteger>
GRACE_HOST_DEVICE uinteger64 space_by_two_21bit(const UInteger x)
{
    return Spread3<uinteger64, 21, 1>::apply(static_cast<uinteger64>(x));
}

} // namespace detail
} // namespace grace
