// grace/cuda/scan.cuh -- segmented scans along per-ray hit lists (reference: cuda/scan.cuh:15-58).
#pragma once
#include "grace/device_vector.h"

namespace grace {

// d_results[i] = sum of d_data[segment start .. i).  d_data and d_results may be the same vector.
template <typename IntVec, typename RealVec>
GRACE_HOST void exclusive_segmented_scan(const IntVec& d_segment_offsets, RealVec& d_data, RealVec& d_results)
{
    static_assert(sizeof(*detail::raw(d_data.data())) == 4, "float data (every reference caller uses float)");
    if (d_results.size() < d_data.size()) d_results.resize(d_data.size());
    GRACE_B200_CHECK(grace_b200_exclusive_segmented_scan_f32(
        detail::context(), detail::raw(d_segment_offsets.data()), d_segment_offsets.size(),
        detail::raw(d_data.data()), d_data.size(), detail::raw(d_results.data()), nullptr));
}

// weighted_values[i] = d_to_sum[i] * d_weights[d_weight_map[i]], then the exclusive segmented scan.
template <typename RealVec, typename MapVec, typename IntVec>
GRACE_HOST void weighted_exclusive_segmented_scan(const RealVec& d_to_sum, const RealVec& d_weights,
                                                  const MapVec& d_weight_map, const IntVec& d_segment_offsets,
                                                  RealVec& d_sum)
{
    static_assert(sizeof(*detail::raw(d_weight_map.data())) == 4, "32-bit weight map");
    if (d_sum.size() < d_to_sum.size()) d_sum.resize(d_to_sum.size());
    GRACE_B200_CHECK(grace_b200_weighted_exclusive_segmented_scan_f32(
        detail::context(), detail::raw(d_to_sum.data()), detail::raw(d_weights.data()),
        (const unsigned*)detail::raw(d_weight_map.data()), detail::raw(d_segment_offsets.data()),
        d_segment_offsets.size(), d_to_sum.size(), detail::raw(d_sum.data()), nullptr));
}

} // namespace grace
