// grace/cuda/kernels/albvh.cuh -- deltas and ALBVH build for arbitrary primitive types (reference:
// cuda/kernels/albvh.cuh:33-47,950-1072).  CUDA only.  The user's delta and AABB
// functors are evaluated by templated kernels here; leaf clustering and the node build run in the
// library (grace_b200_albvh_build_aabb) on the resulting deltas and boxes.
#pragma once
#include "grace/cuda/build_sph.cuh"
#include "grace/cuda/nodes.h"
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/generic/functors/aabb.h"
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
// copy_leaf_deltas (reference: albvh.cuh:51-74): leaf-level deltas, shifted by one like the input.
template <typename DeltaIter, typename LeafDeltaIter>
__global__ void copy_leaf_deltas_kernel(const int4* __restrict__ leaves, const size_t n_leaves, DeltaIter all_deltas,
                                        LeafDeltaIter leaf_deltas)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (tid == 0) leaf_deltas[0] = all_deltas[0];
    for (; tid < n_leaves; tid += stride) {
        const int4 leaf = leaves[tid];
        leaf_deltas[tid + 1] = all_deltas[leaf.x + leaf.y];     // delta after the leaf's last primitive
    }
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
// ---- the build in stages, as the reference's tree-build profilers drive it
//      (tests/profile_tree_gadget/profile_tree_gadget.cu:113-137; reference: albvh.cuh:769-940) ----

// build_leaves: leaves = maximal subtrees with <= max_per_leaf primitives.  The reference writes them
// sparsely (one slot per node, empty ones removed by remove_empty_leaves); here the first L entries of
// d_tmp_leaves are the leaves, dense and in primitive order, and the rest are zeroed, so "leaf.y == 0
// means empty" (cuda/nodes.h:78-88) still holds for anyone looking at the array in between.
// d_tmp_nodes is not needed (kept for source compatibility).  Only the reference's default
// comparison (thrust::less: ties go right) is implemented; delta_comp is not evaluated.
template <typename TmpNodesVec, typename LeavesVec, typename DeltaIter, typename DeltaComp>
GRACE_HOST void build_leaves(TmpNodesVec& /*d_tmp_nodes*/, LeavesVec& d_tmp_leaves, const int max_per_leaf,
                             DeltaIter d_deltas_iter, const DeltaComp /*delta_comp*/)
{
    const size_t n = d_tmp_leaves.size();
    GRACE_CUDA_CHECK(cudaMemsetAsync(detail::raw(d_tmp_leaves.data()), 0, n * sizeof(int4)));
    GRACE_B200_CHECK(grace_b200_albvh_leaves(detail::context(), detail::raw(d_deltas_iter), detail::delta_type(detail::raw(d_deltas_iter)),
                                             n, max_per_leaf, detail::raw(d_tmp_leaves.data()), nullptr, nullptr));
}

// remove_empty_leaves: shrink the tree to the L leaves build_leaves found (albvh.cuh:826-846).
GRACE_HOST void remove_empty_leaves(Tree& d_tree)
{
    int L = 0;
    GRACE_B200_CHECK(grace_b200_albvh_last_n_leaves(detail::context(), &L, nullptr));
    d_tree.nodes.resize(4 * (size_t)(L - 1));
    d_tree.leaves.resize((size_t)L);
}

template <typename LeavesVec, typename DeltaIter, typename LeafDeltaIter>
GRACE_HOST void copy_leaf_deltas(const LeavesVec& d_leaves, DeltaIter d_all_deltas_iter, LeafDeltaIter d_leaf_deltas_iter)
{
    const size_t n = d_leaves.size();
    const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096) + (n == 0);
    ALBVH::copy_leaf_deltas_kernel<<<blocks, 256>>>(detail::raw(d_leaves.data()), n, d_all_deltas_iter, d_leaf_deltas_iter);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
}

// build_nodes: the inner nodes over the (dense) leaves of d_tree, bottom-up (albvh.cuh:854-940).
// Spheres with the reference's AABBSphere functor go straight to the library; any other primitive /
// functor pair has its boxes evaluated here first.  delta_comp as in build_leaves.
template <typename PrimitiveIter, typename DeltaIter, typename DeltaComp, typename AABBFunc>
GRACE_HOST void build_nodes(Tree& d_tree, PrimitiveIter d_prims_iter, DeltaIter d_deltas_iter, const DeltaComp /*delta_comp*/,
                            const AABBFunc AABB)
{
    typedef typename std::remove_cv<typename std::remove_pointer<decltype(detail::raw(d_prims_iter))>::type>::type TPrimitive;
    const size_t L = d_tree.leaves.size();
    const auto* d_deltas = detail::raw(d_deltas_iter);
    if (std::is_same<TPrimitive, float4>::value && std::is_same<AABBFunc, AABBSphere>::value) {
        GRACE_B200_CHECK(grace_b200_albvh_nodes_f4(detail::context(), (const float*)detail::raw(d_prims_iter), d_tree.leaves.data(), L,
                                                   d_deltas, detail::delta_type(d_deltas), d_tree.nodes.data(),
                                                   d_tree.root_index_ptr, nullptr));
        return;
    }
    const int4 last = d_tree.leaves[L - 1];              // leaves are in primitive order: the last one ends the primitives
    const size_t N = (size_t)last.x + last.y;
    device_vector<float4> d_boxes(2 * N);
    const int blocks = (int)((N + 255) / 256 < 4096 ? (N + 255) / 256 : 4096);
    ALBVH::aabbs_kernel<<<blocks, 256>>>(detail::raw(d_prims_iter), N, d_boxes.data(), AABB);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    GRACE_B200_CHECK(grace_b200_albvh_nodes_aabb(detail::context(), (const float*)d_boxes.data(), d_tree.leaves.data(), L, d_deltas,
                                                 detail::delta_type(d_deltas), d_tree.nodes.data(), d_tree.root_index_ptr, nullptr));
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
