// grace/cuda/sort.cuh -- per-ray sort of hit lists (reference: cuda/sort.cuh:100-131).
#pragma once
#include "grace/device_vector.h"

namespace grace {

// Stable ascending sort of every ray's hits by distance; indices and one 32-bit payload
// (e.g. the integrals) are permuted identically.  In place.
template <typename RealVec, typename IntVec, typename IdxVec, typename DataVec>
GRACE_HOST void sort_by_distance(RealVec& d_hit_distances, const IntVec& d_ray_offsets, IdxVec& d_hit_indices,
                                 DataVec& d_hit_data)
{
    static_assert(sizeof(*detail::raw(d_hit_data.data())) == 4, "payload must be 32 bits per hit");
    GRACE_B200_CHECK(grace_b200_sort_by_distance(detail::context(), detail::raw(d_hit_distances.data()),
                                                 detail::raw(d_ray_offsets.data()), d_ray_offsets.size(),
                                                 d_hit_distances.size(), detail::raw(d_hit_indices.data()),
                                                 detail::raw(d_hit_data.data()), nullptr));
}

// Segment index of every element from per-segment offsets (cuda/sort.cuh:20-41).
template <typename IntVec>
GRACE_HOST void offsets_to_segments(const IntVec& d_offsets, IntVec& d_segments)
{
    GRACE_B200_CHECK(grace_b200_offsets_to_segments(detail::context(), detail::raw(d_offsets.data()), d_offsets.size(),
                                                    detail::raw(d_segments.data()), d_segments.size(), nullptr));
}

} // namespace grace
