// grace/cuda/nodes.h -- the tree container (reference: include/grace/cuda/nodes.h:14-58).
//   nodes[4*j+0] = {left child, right child, first leaf, last leaf}   (child >= n_nodes: leaf)
//   nodes[4*j+1] = left box  {bx, tx, by, ty}      nodes[4*j+2] = right box {bx, tx, by, ty}
//   nodes[4*j+3] = {left bz, left tz, right bz, right tz}             (floats bit-cast in int4)
//   leaves[k]    = {first sphere, count, 0, 0}
// As in the reference the constructor sizes the arrays for N_leaves leaves and the builder
// shrinks them to the actual leaf count (albvh.cuh:842-845) -- but the storage is only allocated
// when first touched, and ALBVH_sph touches the node array after it knows the leaf count: 54 MB
// instead of the reference's 1 GiB at 2^24 particles.
#pragma once
#include "grace/device_vector.h"

#include <vector>

namespace grace {

class Tree {
public:
    device_vector<int4> nodes;
    device_vector<int4> leaves;
    int* root_index_ptr;
    int max_per_leaf;

    Tree(size_t N_leaves, int max_per_leaf = 1)
        : nodes(4 * (N_leaves - 1), device_vector<int4>::lazy_t()), leaves(N_leaves, device_vector<int4>::lazy_t()),
          root_index_ptr(nullptr), max_per_leaf(max_per_leaf)
    {
        GRACE_CUDA_CHECK(cudaMalloc((void**)&root_index_ptr, sizeof(int)));
    }
    ~Tree() { cudaFree(root_index_ptr); }
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;
};

// Host-side mirror (reference: cuda/nodes.h:60-76, thrust::host_vector members).
class H_Tree {
public:
    std::vector<int4> nodes;
    std::vector<int4> leaves;
    int root_index;
    int max_per_leaf;

    H_Tree(size_t N_leaves, int _max_per_leaf = 1)
        : nodes(4 * (N_leaves - 1)), leaves(N_leaves), root_index(0), max_per_leaf(_max_per_leaf) {}
    // not in the reference: a copy of a built device tree
    explicit H_Tree(const Tree& d_tree)
        : nodes(d_tree.nodes.to_host()), leaves(d_tree.leaves.to_host()), root_index(0), max_per_leaf(d_tree.max_per_leaf)
    {
        GRACE_CUDA_CHECK(cudaMemcpy(&root_index, d_tree.root_index_ptr, sizeof(int), cudaMemcpyDeviceToHost));
    }
};

// Predicate of the reference's compaction (cuda/nodes.h:78-88): a leaf can never cover zero elements.
struct is_empty_node {
    GRACE_HOST_DEVICE bool operator()(const int4 node) const { return node.y == 0; }
};

namespace detail {
inline grace_b200_tree tree_view(const Tree& t)
{
    grace_b200_tree v;
    v.d_nodes = t.nodes.data();
    v.d_leaves = t.leaves.data();
    v.d_root = t.root_index_ptr;
    v.n_leaves = (int)t.leaves.size();
    v.max_per_leaf = t.max_per_leaf;
    return v;
}
} // namespace detail

} // namespace grace
