// grace/cuda/nodes.h -- the tree container (reference: include/grace/cuda/nodes.h:14-58).
//   nodes[4*j+0] = {left child, right child, first leaf, last leaf}   (child >= n_nodes: leaf)
//   nodes[4*j+1] = left box  {bx, tx, by, ty}      nodes[4*j+2] = right box {bx, tx, by, ty}
//   nodes[4*j+3] = {left bz, left tz, right bz, right tz}             (floats bit-cast in int4)
//   leaves[k]    = {first sphere, count, 0, 0}
// As in the reference the constructor sizes the arrays for N_leaves leaves and the builder
// shrinks them to the actual leaf count (albvh.cuh:842-845).
#pragma once
#include "grace/device_vector.h"

namespace grace {

class Tree {
public:
    device_vector<int4> nodes;
    device_vector<int4> leaves;
    int* root_index_ptr;
    int max_per_leaf;

    Tree(size_t N_leaves, int max_per_leaf = 1)
        : nodes(4 * (N_leaves - 1)), leaves(N_leaves), root_index_ptr(nullptr), max_per_leaf(max_per_leaf)
    {
        GRACE_CUDA_CHECK(cudaMalloc((void**)&root_index_ptr, sizeof(int)));
    }
    ~Tree() { cudaFree(root_index_ptr); }
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;
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
