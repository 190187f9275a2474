// grace/cuda/sort_by_key.cuh -- stable key sort of ARBITRARY records (what the reference's generic
// recipe does with thrust::sort_by_key(keys, primitives), tests/profile_trace_triangle/
// tris_tree.cuh:29).  CUDA only: the library sorts (key, index) pairs, a
// templated kernel here applies the permutation to the user's record type.
#pragma once
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/types.h"

namespace grace {

namespace detail {
template <typename T>
__global__ void gather_kernel(const T* __restrict__ in, T* __restrict__ out, const uinteger32* __restrict__ perm, const size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[perm[i]];
}
inline int sort_perm(uinteger32* k, size_t n, uinteger32* perm) { return grace_b200_sort_pairs_u32(context(), k, nullptr, 0, n, 32, perm, nullptr); }
inline int sort_perm(uinteger64* k, size_t n, uinteger32* perm) { return grace_b200_sort_pairs_u64(context(), k, nullptr, 0, n, 64, perm, nullptr); }
} // namespace detail

// Keys ascending (stable); values permuted identically.  In place.
template <typename KeyType, typename T>
GRACE_HOST void sort_by_key(KeyType* d_keys, const size_t n, T* d_values)
{
    if (n == 0) return;
    device_vector<uinteger32> d_perm(n);
    GRACE_B200_CHECK(detail::sort_perm(d_keys, n, d_perm.data()));
    device_vector<T> d_tmp(n);
    const int blocks = (int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192);
    detail::gather_kernel<<<blocks, 256>>>(d_values, d_tmp.data(), d_perm.data(), n);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    GRACE_CUDA_CHECK(cudaMemcpy(d_values, d_tmp.data(), n * sizeof(T), cudaMemcpyDeviceToDevice));
}

template <typename KeyVec, typename ValueVec>
GRACE_HOST void sort_by_key(KeyVec& d_keys, ValueVec& d_values)
{
    sort_by_key(detail::raw(d_keys.data()), d_keys.size(), detail::raw(d_values.data()));
}

} // namespace grace
