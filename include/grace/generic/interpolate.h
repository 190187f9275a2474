// grace/generic/interpolate.h -- linear table interpolation (reference:
// generic/interpolate.h:11-39).  x in [0, N_table); the result is evaluated in the table's
// precision, as y0 + t*(y1 - y0) (adjacent entries: the difference is exact by Sterbenz).
#pragma once
#include <iterator>
#include "grace/types.h"

namespace grace {

template <typename Real, typename TableIter>
GRACE_HOST_DEVICE Real lerp(Real x, TableIter table, int N_table)
{
    typedef typename std::iterator_traits<TableIter>::value_type TableReal;
    int i = static_cast<int>(x);
    if (i >= N_table - 1) { x = static_cast<TableReal>(N_table - 1); i = N_table - 2; }
    const TableReal y0 = table[i], y1 = table[i + 1];
    const TableReal t = static_cast<TableReal>(x) - i;
#ifdef __CUDA_ARCH__
    return fma(t, y1 - y0, y0);
#else
    return t * (y1 - y0) + y0;
#endif
}

} // namespace grace
