// grace/generic/interpolate.h -- piecewise-linear lookup in an evenly spaced table (reference
// behaviour: generic/interpolate.h:11-39).  `x` is the position in units of the table spacing,
// 0 <= x; positions at or beyond the last entry return the last entry.  The arithmetic runs in
// the TABLE's precision (double for the SPH kernel integrals), whatever Real is:
// y[c] + (x - c) * (y[c+1] - y[c]), fused on the device; adjacent entries are close enough for
// the difference to be exact.
#pragma once
#include <iterator>
#include "grace/types.h"

namespace grace {

template <typename Real, typename TableIter>
GRACE_HOST_DEVICE Real lerp(Real x, TableIter table, int N_table)
{
    typedef typename std::iterator_traits<TableIter>::value_type Y;
    const int last = N_table - 1;
    int cell = static_cast<int>(x);
    Y pos = static_cast<Y>(x);
    if (cell >= last) {               // clamp to the right edge of the last cell
        cell = last - 1;
        pos = static_cast<Y>(last);
    }
    const Y left = table[cell];
    const Y rise = table[cell + 1] - left;
    const Y frac = pos - cell;
#ifdef __CUDA_ARCH__
    return fma(frac, rise, left);
#else
    return frac * rise + left;
#endif
}

} // namespace grace
