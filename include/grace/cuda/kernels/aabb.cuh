// grace/cuda/kernels/aabb.cuh -- centroids of arbitrary primitives through the user's functor
// (reference: include/grace/cuda/kernels/aabb.cuh:14-49; the tree-build profilers call it before
// min_vec3/max_vec3, tests/profile_tree_gadget/profile_tree_gadget.cu:91-98).  A user functor cannot
// cross the C ABI, so this stays a header template instantiated in the caller's translation unit.
#pragma once
#include "grace/cuda/kernel_config.h"
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/generic/functors/aabb.h"

#include <iterator>

namespace grace {
namespace AABB {

template <typename PrimitiveIter, typename CentroidIter, typename CentroidFunc>
__global__ void compute_centroids_kernel(PrimitiveIter primitives, const size_t N_primitives, CentroidIter centroids,
                                         const CentroidFunc centroid)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N_primitives; i += (size_t)gridDim.x * blockDim.x)
        centroids[i] = centroid(primitives[i]);
}

template <typename PrimitiveIter, typename CentroidIter, typename CentroidFunc>
GRACE_HOST void compute_centroids(PrimitiveIter d_prims_iter, const size_t N_primitives, CentroidIter d_centroid_iter,
                                  const CentroidFunc centroid)
{
    if (N_primitives == 0) return;
    const size_t want = (N_primitives + 255) / 256;
    const int blocks = (int)(want < (size_t)MAX_BLOCKS ? want : (size_t)MAX_BLOCKS);
    compute_centroids_kernel<<<blocks, 256>>>(d_prims_iter, N_primitives, d_centroid_iter, centroid);
    GRACE_KERNEL_CHECK();
}

} // namespace AABB
} // namespace grace
