// grace/cuda/kernels/bintree_trace.cuh -- traversal with user-defined primitive types and
// functors (reference: cuda/kernels/bintree_trace.cuh:52-367).  CUDA only.
//
// Templates over user functors are instantiated in the USER's translation unit and cannot cross
// the C ABI, so this header carries the kernel itself.  It follows the reference's packet schedule
// exactly, because user functors may have side effects that depend on what is tested and in
// which order (e.g. a closest-hit functor narrowing t_min): 32 consecutive rays share one
// traversal; a child is entered when ANY ray's slab test reaches it (the reference's slab
// arithmetic bit for bit, device/intersect.cuh), right child pushed first so the left subtree
// is walked first; at a leaf EVERY lane calls Intersection for EVERY primitive in index order.
// What changes is the machinery around it: packets are handed to a persistent grid sized to
// the device by an atomic ticket (the reference caps the grid at 112 blocks), the top of the
// stack lives in a register, the node is fetched with four 128-bit loads issued together and
// leaf primitives are staged once per packet.
#pragma once
#include "grace/cuda/device/intersect.cuh"
#include "grace/cuda/functors/trace.cuh"
#include "grace/cuda/nodes.h"
#include "grace/device_vector.h"
#include "grace/error.h"
#include "grace/ray.h"

#include <iterator>
#include <stdexcept>

namespace grace {

namespace gpu {

constexpr int GENERIC_TRACE_THREADS = 128;
constexpr int GENERIC_STACK_SIZE = 96;      // reference STACK_SIZE is 64 (kernel_config.h:13)

template <typename RayData, typename TPrimitive, typename Init, typename Intersection, typename OnHit,
          typename OnRayEntry, typename OnRayExit>
__global__ void __launch_bounds__(GENERIC_TRACE_THREADS)
trace_kernel(const Ray* __restrict__ rays, const int n_packets, const int4* __restrict__ nodes, const int n_nodes,
             const int4* __restrict__ leaves, const int* __restrict__ root_index, const TPrimitive* __restrict__ primitives,
             const int max_per_leaf, const size_t user_smem_bytes, int* packet_counter, int* overflow_flag,
             Init init, Intersection intersect, OnHit on_hit, OnRayEntry ray_entry, OnRayExit ray_exit)
{
    extern __shared__ __align__(16) char smem_trace[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int N_warps = GENERIC_TRACE_THREADS / 32;
    const BoundIter<char> sm_iter_usr(smem_trace, user_smem_bytes);
    init(sm_iter_usr);
    __syncthreads();

    // layout: [user block | pad to 16 | stacks | primitives]
    size_t off = (user_smem_bytes + 15) & ~(size_t)15;
    int* stack = reinterpret_cast<int*>(smem_trace + off) + wid * GENERIC_STACK_SIZE;
    off += (size_t)N_warps * GENERIC_STACK_SIZE * sizeof(int);
    const size_t palign = alignof(TPrimitive) > 16 ? alignof(TPrimitive) : 16;
    off = (off + palign - 1) / palign * palign;
    TPrimitive* sm_prims = reinterpret_cast<TPrimitive*>(smem_trace + off) + (size_t)wid * max_per_leaf;
    const int root = *root_index;

    for (;;) {
        int packet = 0;
        if (lane == 0) packet = atomicAdd(packet_counter, 1);
        packet = __shfl_sync(0xffffffffu, packet, 0);
        if (packet >= n_packets) break;
        const int ray_index = packet * 32 + lane;
        const Ray ray = rays[ray_index];
        RayData ray_data = {};
        ray_entry(ray_index, ray, ray_data, sm_iter_usr);
        float3 invd, origin;
        invd.x = 1.f / ray.dx; invd.y = 1.f / ray.dy; invd.z = 1.f / ray.dz;
        origin.x = ray.ox; origin.y = ray.oy; origin.z = ray.oz;

        int sp = 0;            // entries below the register top
        int top = root;        // -1: empty
        while (top >= 0) {
            if (top < n_nodes) {
                const int4* np = nodes + 4 * (size_t)top;
                const int4 node = __ldg(np + 0);
                const int4 l = __ldg(np + 1), r = __ldg(np + 2), lr = __ldg(np + 3);
                const int lr_hit = AABBs_hit(invd, origin, ray.length,
                    make_float4(__int_as_float(l.x), __int_as_float(l.y), __int_as_float(l.z), __int_as_float(l.w)),
                    make_float4(__int_as_float(r.x), __int_as_float(r.y), __int_as_float(r.z), __int_as_float(r.w)),
                    make_float4(__int_as_float(lr.x), __int_as_float(lr.y), __int_as_float(lr.z), __int_as_float(lr.w)));
                const bool any_r = __any_sync(0xffffffffu, lr_hit & 1);
                const bool any_l = __any_sync(0xffffffffu, lr_hit >= 2);
                // pop, push right, push left: the left child ends up on top
                if (any_l && any_r) {
                    if (sp >= GENERIC_STACK_SIZE) { if (lane == 0) *overflow_flag = 1; top = -1; sp = 0; continue; }
                    stack[sp++] = node.y;
                    top = node.x;
                } else if (any_l) {
                    top = node.x;
                } else if (any_r) {
                    top = node.y;
                } else {
                    top = sp > 0 ? stack[--sp] : -1;
                }
            } else {
                const int4 leaf = __ldg(leaves + (top - n_nodes));
                top = sp > 0 ? stack[--sp] : -1;
                __syncwarp();
                for (int i = lane; i < leaf.y; i += 32) sm_prims[i] = primitives[leaf.x + i];
                __syncwarp();
                for (int i = 0; i < leaf.y; ++i) {
                    const TPrimitive prim = sm_prims[i];
                    if (intersect(ray, prim, ray_data, i, sm_iter_usr))
                        on_hit(ray_index, ray, ray_data, leaf.x + i, prim, i, sm_iter_usr);
                }
            }
        }
        ray_exit(ray_index, ray, ray_data, sm_iter_usr);
    }
}

} // namespace gpu

// trace<RayData>(rays, n_rays, primitives, n_primitives, tree, user_smem_bytes, init, intersect,
//                on_hit, ray_entry, ray_exit)   -- reference: bintree_trace.cuh:208-285
template <typename RayData, typename TPrimitive, typename Init, typename Intersection, typename OnHit,
          typename OnRayEntry, typename OnRayExit>
GRACE_HOST void trace(const Ray* d_rays, const size_t N_rays, const TPrimitive* d_primitives, const size_t /*N_primitives*/,
                      const Tree& d_tree, const size_t user_smem_bytes, Init init, Intersection intersect, OnHit on_hit,
                      OnRayEntry ray_entry, OnRayExit ray_exit)
{
    if (N_rays % 32 != 0)     // bintree_trace.cuh:231-238
        throw std::invalid_argument("Number of rays must be a multiple of the warp size (32).");
    if (N_rays == 0) return;
    const size_t n_leaves = d_tree.leaves.size();
    auto kernel = gpu::trace_kernel<RayData, TPrimitive, Init, Intersection, OnHit, OnRayEntry, OnRayExit>;
    const size_t palign = alignof(TPrimitive) > 16 ? alignof(TPrimitive) : 16;
    const size_t smem = ((user_smem_bytes + 15) & ~(size_t)15) + (gpu::GENERIC_TRACE_THREADS / 32) * gpu::GENERIC_STACK_SIZE * sizeof(int)
                        + palign + (gpu::GENERIC_TRACE_THREADS / 32) * (size_t)d_tree.max_per_leaf * sizeof(TPrimitive);
    GRACE_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    GRACE_CUDA_CHECK(cudaGetDevice(&dev));
    GRACE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GRACE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, gpu::GENERIC_TRACE_THREADS, smem));
    const int n_packets = (int)(N_rays / 32);
    int blocks = sms * (per_sm > 0 ? per_sm : 1);
    const int need = (n_packets + gpu::GENERIC_TRACE_THREADS / 32 - 1) / (gpu::GENERIC_TRACE_THREADS / 32);
    if (blocks > need) blocks = need;
    device_vector<int> d_scalars(std::vector<int>(2, 0));      // packet ticket, stack-overflow flag
    kernel<<<blocks, gpu::GENERIC_TRACE_THREADS, smem>>>(
        d_rays, n_packets, d_tree.nodes.data(), (int)n_leaves - 1, d_tree.leaves.data(), d_tree.root_index_ptr,
        d_primitives, d_tree.max_per_leaf, user_smem_bytes, d_scalars.data(), d_scalars.data() + 1, init, intersect,
        on_hit, ray_entry, ray_exit);
    GRACE_CUDA_CHECK(cudaPeekAtLastError());
    // releasing d_scalars waits for the kernel anyway: look at the stack-overflow flag first, results of an
    // overflowed walk are short (the reference asserts in the kernel, debug builds only: bintree_trace.cuh:162-164)
    if (d_scalars[1] != 0)
        throw std::runtime_error("grace::trace: traversal stack overflow (tree deeper than GENERIC_STACK_SIZE allows)");
}

template <typename RayData, typename RayVec, typename PrimVec, typename Init, typename Intersection, typename OnHit,
          typename OnRayEntry, typename OnRayExit>
GRACE_HOST void trace(const RayVec& d_rays, const PrimVec& d_primitives, const Tree& d_tree, const size_t user_smem_bytes,
                      Init init, Intersection intersect, OnHit on_hit, OnRayEntry ray_entry, OnRayExit ray_exit)
{
    trace<RayData>(detail::raw(d_rays.data()), d_rays.size(), detail::raw(d_primitives.data()), d_primitives.size(), d_tree,
                   user_smem_bytes, init, intersect, on_hit, ray_entry, ray_exit);
}

// The reference's texture-reference variant (bintree_trace.cuh:290-367); texture references no
// longer exist in CUDA, the read-only data path serves the same purpose.
template <typename RayData, typename RayVec, typename PrimVec, typename Init, typename Intersection, typename OnHit,
          typename OnRayEntry, typename OnRayExit>
GRACE_HOST void trace_texref(const RayVec& d_rays, const PrimVec& d_primitives, const Tree& d_tree,
                             const size_t user_smem_bytes, Init init, Intersection intersect, OnHit on_hit,
                             OnRayEntry ray_entry, OnRayExit ray_exit)
{
    trace<RayData>(d_rays, d_primitives, d_tree, user_smem_bytes, init, intersect, on_hit, ray_entry, ray_exit);
}

} // namespace grace
