// grace/cuda/kernels/albvh.cuh -- deltas and ALBVH build for arbitrary primitive types (reference:
// cuda/kernels/albvh.cuh:33-47,950-1072).  CUDA only.  The user's delta and AABB
// functors are evaluated by templated kernels here; leaf clustering and the node build run in the
// library (grace_b200_albvh_build_aabb) on the resulting deltas and boxes.
#pragma once
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/nodes.h"
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/generic/functors/albvh.h"

#include <iterator>

namespace grace {

namespace ALBVH {
template <typename KeyIter, typename DeltaIter, typename DeltaFunc>
__global__ void compute_deltas_kernel(KeyIter keys, const size_t n_keys, DeltaIter deltas, const DeltaFunc delta_func)
{
    // the range [-1, n_keys) is valid for delta_func; deltas is shifted by one
    for (size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tid <= n_keys; tid += (size_t)gridDim.x * blockDim.x)
        deltas[tid] = delta_func((int)tid - 1, keys, n_keys);
}
template <typename TPrimitive, typename AABBFunc>
__global__ void aabbs_kernel(const TPrimitive* __restrict__ prims, const size_t n, float4* __restrict__ boxes, const AABBFunc AABB)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float3 bot, top;
        AABB(prims[i], &bot, &top);
        boxes[2 * i] = make_float4(bot.x, bot.y, bot.z, 0.f);
        boxes[2 * i + 1] = make_float4(top.x, top.y, top.z, 0.f);
    }
}
} // namespace ALBVH

template <typename KeyIter, typename DeltaIter, typename DeltaFunc>
GRACE_HOST void compute_deltas(KeyIter d_keys_iter, const size_t N_keys, DeltaIter d_deltas_iter, const DeltaFunc delta_func)
{
    const size_t n = N_keys + 1;
    const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    ALBVH::compute_deltas_kernel<<<blocks, 256>>>(d_keys_iter, N_keys, d_deltas_iter, delta_func);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
}

template <typename KeyVec, typename DeltaVec, typename DeltaFunc,
          typename = decltype(std::declval<const KeyVec&>().size())>
GRACE_HOST void compute_deltas(const KeyVec& d_keys, DeltaVec& d_deltas, const DeltaFunc delta_func)
{
    compute_deltas(detail::raw(d_keys.data()), d_keys.size(), detail::raw(d_deltas.data()), delta_func);
}

// build_ALBVH(tree, primitives, N, deltas, AABBFunc[, wipe]) with the reference's default comparison
// (thrust::less: ties go right).  `wipe` is accepted for source compatibility; the builder writes
// every element it defines.
template <typename TPrimitive, typename DeltaType, typename AABBFunc>
GRACE_HOST void build_ALBVH(Tree& d_tree, const TPrimitive* d_prims, const size_t N_primitives, const DeltaType* d_deltas,
                            const AABBFunc AABB, const bool /*wipe*/ = false)
{
    device_vector<float4> d_boxes(2 * N_primitives);
    const int blocks = (int)((N_primitives + 255) / 256 < 4096 ? (N_primitives + 255) / 256 : 4096);
    ALBVH::aabbs_kernel<<<blocks, 256>>>(d_prims, N_primitives, d_boxes.data(), AABB);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    int L = 0;
    GRACE_B200_CHECK(grace_b200_albvh_build_aabb(detail::context(), (const float*)d_boxes.data(), N_primitives, d_deltas,
                                                 detail::delta_type(d_deltas), d_tree.max_per_leaf, d_tree.nodes.data(),
                                                 d_tree.leaves.data(), d_tree.root_index_ptr, &L, nullptr));
    d_tree.nodes.resize(4 * (size_t)(L - 1));      // remove_empty_leaves, albvh.cuh:842-845
    d_tree.leaves.resize((size_t)L);
}

template <typename PrimVec, typename DeltaVec, typename AABBFunc,
          typename = decltype(std::declval<const PrimVec&>().size())>
GRACE_HOST void build_ALBVH(Tree& d_tree, const PrimVec& d_primitives, const DeltaVec& d_deltas, const AABBFunc AABB,
                            const bool wipe = false)
{
    build_ALBVH(d_tree, detail::raw(d_primitives.data()), d_primitives.size(), detail::raw(d_deltas.data()), AABB, wipe);
}

} // namespace grace
