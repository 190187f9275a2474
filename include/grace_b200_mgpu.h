/*
 * grace_b200_mgpu.h -- C ABI of the multi-GPU layer of the B200-native GRACE hot path
 * (libgrace_b200_mgpu.so = libgrace_b200.so + NCCL).
 *
 * GRACE itself has no multi-GPU code: its profilers take one device id
 * (tests/profile_one_to_many_rays_gadget/profile_one_to_many_rays_gadget.cu:43-52).  Rays are
 * independent given the tree, so the layer is: ONE host process driving every device of the node
 * (ncclCommInitAll), the sorted particles and the tree REPLICATED on every device, the rays dealt
 * round-robin in 32-aligned tiles of 4096, each device tracing its tiles, and the 4 bytes/ray
 * results gathered to device 0 and put back in ray order.  Packets are 32 consecutive rays
 * (include/grace/cuda/kernels/bintree_trace.cuh:75,231-238) and tile boundaries are multiples of
 * 32, so every ray shares its packet with the same neighbours as in a one-GPU run: results are
 * bit-identical for every device count.  NCCL is used only to broadcast (particles, or the
 * finished tree) and to gather the per-ray outputs.
 *
 * All pointers named h_* are HOST memory (pinned memory makes the copies asynchronous).  Every
 * call returns when its results are complete.  Return value: a GRACE_B200_* code;
 * grace_b200_mgpu_last_error() has the message.  One handle must not be used from two threads at once.
 */
#ifndef GRACE_B200_MGPU_H
#define GRACE_B200_MGPU_H

#include "grace_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct grace_b200_mgpu grace_b200_mgpu;

/* How the replicated tree comes about. */
#define GRACE_B200_MGPU_BUILD_EVERYWHERE 0   /* broadcast the particles, every device builds (deterministic: same bits) */
#define GRACE_B200_MGPU_BUILD_ON_ROOT    1   /* device 0 builds, the sorted particles + nodes + leaves + root are broadcast */

/* n_devices CUDA devices: `devices[i]`, or 0..n_devices-1 when devices is NULL; n_devices <= 0 means
 * all visible devices.  Creates one grace_b200 context, one stream and one NCCL rank per device. */
int grace_b200_mgpu_init(grace_b200_mgpu** mg, int n_devices, const int* devices);
int grace_b200_mgpu_finalize(grace_b200_mgpu* mg);
int grace_b200_mgpu_n_devices(const grace_b200_mgpu* mg);
const char* grace_b200_mgpu_last_error(void);

/* Particles (host, n x float4 {x,y,z,h}, in any order) -> sorted particles and ALBVH on every device:
 * the tests/helper/tree.cuh:15-43 recipe (30- or 63-bit keys + sort, Euclidean deltas, ALBVH).
 * *h_n_leaves (may be NULL) receives the leaf count.  ms3 (may be NULL) receives the device-0 times in
 * ms of {host->device copy, broadcast(s), build}. */
int grace_b200_mgpu_build_f4(grace_b200_mgpu* mg, const float* h_spheres4, size_t n, int max_per_leaf,
                             int key_bits, int how, int* h_n_leaves, float* ms3);

/* replaces: trace_cumulative_sph / trace_hitcounts_sph (cuda/trace_sph.cuh:58-109) over all devices.
 * h_rays: n_rays x grace::Ray on the host (n_rays % 32 == 0); h_out: n_rays results in ray order.
 * ms4 (may be NULL): {rays to the devices, trace (max over devices), gather + reassembly, result to host}. */
int grace_b200_mgpu_trace_cumulative_f4(grace_b200_mgpu* mg, const grace_b200_ray* h_rays, size_t n_rays,
                                        float* h_cumulated, float* ms4);
int grace_b200_mgpu_trace_hitcounts_f4(grace_b200_mgpu* mg, const grace_b200_ray* h_rays, size_t n_rays,
                                       int* h_hit_counts, float* ms4);

/* Diagnostics for the tests: copies of device `dev`'s sorted particles / nodes to the host. */
int grace_b200_mgpu_copy_tree(grace_b200_mgpu* mg, int dev, float* h_spheres4, int* h_nodes16, int* h_leaves4, int* h_root);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GRACE_B200_MGPU_H */
