// grace/cuda/sort.cuh -- per-ray sort of hit lists (reference: cuda/sort.cuh:100-131).
#pragma once
#include "grace/device_vector.h"

// The reference's sort.cuh pulls in sgpu, and with it the driver API header: user code written
// against GRACE calls cuMemGetInfo() without including <cuda.h> itself
// (tests/profile_trace_gadget/profile_trace_gadget.cu:160).
#include <cuda.h>

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

#ifdef __CUDACC__
namespace detail {
template <typename IndexType, typename T>
__global__ void gather_kernel(const IndexType* __restrict__ idx, const T* __restrict__ src, T* __restrict__ dst, const size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
} // namespace detail

// d_unordered[i] <- d_unordered[d_indices[i]] (reference: cuda/sort.cuh:43-51, a copy + thrust::gather).
template <typename IdxVec, typename Vec>
GRACE_HOST void order_by_index(const IdxVec& d_indices, Vec& d_unordered)
{
    typedef typename detail::elem_of<Vec>::type T;
    const size_t n = d_indices.size();
    if (n == 0) return;
    device_vector<T> d_tmp(d_unordered.size());
    GRACE_CUDA_CHECK(cudaMemcpyAsync(d_tmp.data(), detail::raw(d_unordered.data()), d_unordered.size() * sizeof(T), cudaMemcpyDeviceToDevice));
    const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    detail::gather_kernel<<<blocks, 256>>>(detail::raw(d_indices.data()), d_tmp.data(), detail::raw(d_unordered.data()), n);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    GRACE_CUDA_CHECK(cudaDeviceSynchronize());      // d_tmp is released on return
}
#endif // __CUDACC__

// Segment index of every element from per-segment offsets (cuda/sort.cuh:20-41).
template <typename IntVec>
GRACE_HOST void offsets_to_segments(const IntVec& d_offsets, IntVec& d_segments)
{
    GRACE_B200_CHECK(grace_b200_offsets_to_segments(detail::context(), detail::raw(d_offsets.data()), d_offsets.size(),
                                                    detail::raw(d_segments.data()), d_segments.size(), nullptr));
}

} // namespace grace
