/*
 * grace_b200.h -- C ABI of the B200-native GRACE ray-tracing hot path.
 *
 * GRACE (spthm/grace-devel) has no C ABI: its boundary is a set of header
 * templates in namespace grace that user .cu files call directly.  Each entry
 * point below replaces one of those templates for the SPH (float4 sphere) path;
 * the citation after "replaces:" is the reference interface (paths relative to
 * the GRACE tree).  include/grace/ *.h in this repo re-creates the reference's
 * C++ names on top of this ABI.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; h_* is host
 *     memory; sizes are element counts;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL is
 *     the legacy default stream).  Calls that must return a count to the host
 *     (documented per function) synchronise that stream, the others do not;
 *   - temporaries come from the context's workspace arena, which only grows, and from a block
 *     of device scalars (tickets, counters, the traversal's error flag): ONE STREAM PER
 *     CONTEXT AT A TIME.  Calls on the same context are ordered by enqueueing them on the same
 *     stream; work on two streams (or from two host threads) needs two contexts, or the caller
 *     must order the second stream behind the first (event) before its next call.  Workspace
 *     growth synchronises the device: pre-size it (grace_b200_reserve) before capturing a
 *     CUDA graph.  (The reference is not re-entrant either: cuda/kernels/bintree_trace.cuh:37-38);
 *   - return value: 0 = success, otherwise a GRACE_B200_E* code;
 *     grace_b200_last_error() gives the message of the calling thread's last
 *     failure.  The C++ shim maps GRACE_B200_EINVAL to std::invalid_argument and
 *     CUDA failures to print+exit, as the reference does (error.h:35-64,
 *     bintree_trace.cuh:231-238, albvh.cuh:795-799).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef GRACE_B200_H
#define GRACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define GRACE_B200_OK       0
#define GRACE_B200_EINVAL   1   /* bad argument (reference: std::invalid_argument) */
#define GRACE_B200_ECUDA    2   /* CUDA runtime failure */
#define GRACE_B200_ERANGE   3   /* size exceeds what the 32-bit reference layout can hold */
#define GRACE_B200_ENOMEM   4
#define GRACE_B200_EDEVICE  5   /* device-side traversal failure (stack overflow, malformed tree): outputs incomplete */

/* Delta (key-difference) element types accepted by the tree builder. */
#define GRACE_B200_DELTA_F32 0
#define GRACE_B200_DELTA_U32 1
#define GRACE_B200_DELTA_U64 2

/* grace::RaySortType, include/grace/types.h:47-51 */
#define GRACE_B200_NO_SORT        0
#define GRACE_B200_DIRECTION_SORT 1
#define GRACE_B200_ENDPOINT_SORT  2

typedef struct grace_b200_ctx grace_b200_ctx;

/* grace::Ray, include/grace/ray.h:5-10 (7 floats, 28 bytes, direction normalised). */
typedef struct { float dx, dy, dz, ox, oy, oz, length; } grace_b200_ray;

/* ---- context ---------------------------------------------------------------- */
int grace_b200_create(grace_b200_ctx** ctx, int device);
int grace_b200_destroy(grace_b200_ctx* ctx);
const char* grace_b200_last_error(void);
const char* grace_b200_version(void);
/* Pre-size the workspace arena (optional; avoids growth inside timed regions). */
int grace_b200_reserve(grace_b200_ctx* ctx, size_t bytes);
size_t grace_b200_workspace_bytes(const grace_b200_ctx* ctx);

/* ---- bounds + Morton keys --------------------------------------------------- */
/* replaces: AABB::compute_centroids + min_vec3/max_vec3
 *   (cuda/kernels/aabb.cuh:14-49, cuda/util/extrema.cuh:502-513,667-678) as used by
 *   morton_keys(prims, keys, centroid, bots, tops) (cuda/kernels/morton.cuh:139-173).
 * d_bounds6 receives {min x,y,z, max x,y,z} of the sphere centres (device). */
int grace_b200_bounds_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                         float* d_bounds6, void* stream);
/* Component-wise min and max of all four components {x,y,z,w} (min_vec4/max_vec4,
 * min_max_x: cuda/util/extrema.cuh:189-230,456-731).  d_minmax8 = {min xyzw, max xyzw}. */
int grace_b200_minmax_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                         float* d_minmax8, void* stream);

/* The same, returned to the host (h_minmax8; the stream is synchronised): what the reference's
 * min_vec3 / max_vec3 / min_max_x hand back (cuda/util/extrema.cuh:189-230,502-513). */
int grace_b200_minmax_f4_host(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                              float* h_minmax8, void* stream);

/* replaces: morton_keys_sph / morton::morton_keys_kernel
 *   (cuda/build_sph.cuh:19-34, cuda/kernels/morton.cuh:30-55,97-116).
 * key = interleave(trunc(scale*(c - bot))), scale = span/(top-bot), 10 or 21 bits
 * per axis.  Bounds are read from device memory so no host sync is needed. */
int grace_b200_morton_keys30_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, uint32_t* d_keys, void* stream);
int grace_b200_morton_keys63_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, uint64_t* d_keys, void* stream);

/* ---- stable key-value radix sort -------------------------------------------- */
/* replaces: thrust::sort_by_key(keys, values) (cuda/build_sph.cuh:46,57,70,81;
 *   cuda/kernels/gen_rays.cuh:483,520,577,615).  Stable LSD onesweep sort on key
 * bits [0, key_bits).  Keys and values are sorted in place; value_bytes is the
 * record size (16 = float4 sphere, 28 = Ray, 4 = plain 32-bit payload).
 * d_perm (optional, may be NULL) receives the permutation sorted[i] = in[perm[i]]. */
int grace_b200_sort_pairs_u32(grace_b200_ctx* ctx, uint32_t* d_keys, void* d_values,
                              int value_bytes, size_t n, int key_bits,
                              uint32_t* d_perm, void* stream);
int grace_b200_sort_pairs_u64(grace_b200_ctx* ctx, uint64_t* d_keys, void* d_values,
                              int value_bytes, size_t n, int key_bits,
                              uint32_t* d_perm, void* stream);

/* replaces: morton_keys30_sort_sph / morton_keys63_sort_sph (cuda/build_sph.cuh:41-82).
 * key_bits = 30 or 63.  h_bot3/h_top3 = explicit bounds, or both NULL to compute
 * them from the centres.  d_keys_out (optional) receives the sorted keys
 * (uint32_t[n] or uint64_t[n]).  Spheres are sorted in place. */
int grace_b200_morton_sort_f4(grace_b200_ctx* ctx, float* d_spheres4, size_t n, int key_bits,
                              const float* h_bot3, const float* h_top3,
                              void* d_keys_out, void* stream);

/* ---- deltas ----------------------------------------------------------------- */
/* replaces: compute_deltas + DeltaEuclidean / DeltaSurfaceArea / DeltaXOR
 *   (cuda/kernels/albvh.cuh:33-47,950-978; generic/functors/albvh.h:17-126;
 *    cuda/build_sph.cuh:87-114).  Outputs n+1 deltas, shifted by one, with
 *   +inf / all-ones sentinels at both ends. */
int grace_b200_deltas_euclid_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                float* d_deltas, void* stream);
int grace_b200_deltas_sarea_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                               float* d_deltas, void* stream);
int grace_b200_deltas_xor32(grace_b200_ctx* ctx, const uint32_t* d_keys, size_t n,
                            uint32_t* d_deltas, void* stream);
int grace_b200_deltas_xor64(grace_b200_ctx* ctx, const uint64_t* d_keys, size_t n,
                            uint64_t* d_deltas, void* stream);

/* ---- ALBVH build ------------------------------------------------------------ */
/* replaces: build_ALBVH / ALBVH_sph (cuda/kernels/albvh.cuh:986-1072,
 *   cuda/build_sph.cuh:118-124) = build_leaves + remove_empty_leaves +
 *   copy_leaf_deltas + build_nodes.
 * d_nodes  : int4[4*(n-1)] capacity; on return the first 4*(L-1) hold the
 *            reference layout (cuda/nodes.h:21-36);
 * d_leaves : int4[n] capacity; first L valid, {first, count, 0, 0};
 * d_root   : device int, index of the root node (Tree::root_index_ptr);
 * h_n_leaves: host int receiving L; when non-NULL the stream is synchronised.
 *            When NULL nothing is synchronised and L can be read later with
 *            grace_b200_albvh_last_n_leaves().
 * Returns GRACE_B200_EINVAL if n <= max_per_leaf (albvh.cuh:795-799). */
int grace_b200_albvh_build_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                              const void* d_deltas, int delta_type, int max_per_leaf,
                              void* d_nodes, void* d_leaves, int* d_root,
                              int* h_n_leaves, void* stream);
int grace_b200_albvh_last_n_leaves(grace_b200_ctx* ctx, int* h_n_leaves, void* stream);
/* replaces: build_ALBVH(tree, primitives, deltas, AABBFunc) for arbitrary primitive types
 * (cuda/kernels/albvh.cuh:986-1072): the caller evaluates its AABB functor
 * (include/grace/cuda/kernels/albvh.cuh here does) into d_aabbs8 = two float4 per primitive,
 * {bx,by,bz,-} {tx,ty,tz,-}; everything else as grace_b200_albvh_build_f4. */
int grace_b200_albvh_build_aabb(grace_b200_ctx* ctx, const float* d_aabbs8, size_t n,
                                const void* d_deltas, int delta_type, int max_per_leaf,
                                void* d_nodes, void* d_leaves, int* d_root, int* h_n_leaves,
                                void* stream);

/* The build in stages, as the reference's tree-build profilers time it
 * (tests/profile_tree_gadget/profile_tree_gadget.cu:113-137, tests/profile_tree/profile_tree.cu:113-133).
 * replaces: ALBVH::build_leaves + the compaction of ALBVH::remove_empty_leaves
 *   (cuda/kernels/albvh.cuh:785-846).  d_leaves: int4[n] capacity; on return the first L entries are
 *   the leaves {first, count, 0, 0}, dense and ordered by primitive range (what remove_if leaves
 *   behind); entries past L are not written.  L goes to *h_n_leaves (stream synchronised) or, with
 *   NULL, stays on the device for grace_b200_albvh_last_n_leaves(). */
int grace_b200_albvh_leaves(grace_b200_ctx* ctx, const void* d_deltas, int delta_type, size_t n,
                            int max_per_leaf, void* d_leaves, int* h_n_leaves, void* stream);
/* replaces: ALBVH::build_nodes (cuda/kernels/albvh.cuh:854-940).  d_leaf_deltas: n_leaves + 1 deltas
 *   between consecutive leaves, shifted by one, sentinels at both ends (what ALBVH::copy_leaf_deltas,
 *   albvh.cuh:51-74,769-782, extracts); d_nodes: int4[4*(n_leaves-1)].  _aabb: d_aabbs8 as in
 *   grace_b200_albvh_build_aabb.  d_leaf_deltas may be NULL when the call directly follows
 *   grace_b200_albvh_leaves on the same context (no other grace_b200 call in between): the leaf-level
 *   deltas that call extracted are still in the workspace.  This is how the one-call build sizes the
 *   node array from the leaf count instead of from the primitive count. */
int grace_b200_albvh_nodes_f4(grace_b200_ctx* ctx, const float* d_spheres4, const void* d_leaves,
                              size_t n_leaves, const void* d_leaf_deltas, int delta_type,
                              void* d_nodes, int* d_root, void* stream);
int grace_b200_albvh_nodes_aabb(grace_b200_ctx* ctx, const float* d_aabbs8, const void* d_leaves,
                                size_t n_leaves, const void* d_leaf_deltas, int delta_type,
                                void* d_nodes, int* d_root, void* stream);

/* ---- trace ------------------------------------------------------------------ */
/* A tree as the trace entry points take it (grace::Tree, cuda/nodes.h:14-58). */
typedef struct {
    const void* d_nodes;    /* int4[4*(n_leaves-1)] */
    const void* d_leaves;   /* int4[n_leaves]       */
    const int*  d_root;     /* device int           */
    int n_leaves;
    int max_per_leaf;
} grace_b200_tree;

/* Traversal schedule (all three return identical results wherever the reference passes
 * its own brute-force test, tests/tree_traversal/tree_traversal.cu:84-121).
 *  PACKET (default): 32 consecutive rays share one traversal as in the reference
 *     (cuda/kernels/bintree_trace.cuh:119-193), with a conservatively padded slab test,
 *     staged leaves and deferred on-hit work; the hit set is exactly the brute-force set.
 *  PER_RAY: every lane walks the tree for its own ray (padded slab test).
 *  PACKET_REF: the reference's schedule and slab arithmetic bit for bit; defines the
 *     traversal counters of grace_b200_trace_stats_f4. */
#define GRACE_B200_TRACE_PER_RAY    0
#define GRACE_B200_TRACE_PACKET     1
#define GRACE_B200_TRACE_PACKET_REF 2
int grace_b200_set_trace_mode(grace_b200_ctx* ctx, int mode);
/* Load balancing of the PACKET schedule.  Hit counts and column densities: work stealing inside
 * the launch -- a warp without a packet asks the longest-running unit for the bottom subtree of
 * its stack and walks it for all rays of that unit (the terms of a ray are still added in
 * ascending primitive order, so results are unaffected); a unit can be robbed once it has run
 * `steps` inner-node + leaf visits.  Hit lists: once every packet of a launch has been claimed, a
 * packet still running after `steps` visits is suspended and resumed as tasks over disjoint
 * subsets of its rays.  0 disables both; the default is 16 for work stealing and 64 for the hit-list rounds.  OR-ing GRACE_B200_BUDGET_EAGER into
 * `steps` makes any subtree worth stealing and suspends every hit-list unit at `steps` whether or
 * not unclaimed work is left (used by the tests to force splits). */
#define GRACE_B200_BUDGET_EAGER (1 << 30)
int grace_b200_set_trace_budget(grace_b200_ctx* ctx, int steps);
/* (Also the pool in which grace_b200_trace_hits_count_f4 records the hits themselves, 16 bytes each, so that
 * grace_b200_trace_hits_fill_f4 needs no second traversal; default there: 32 KiB per ray, 256 MiB to 4 GiB,
 * and a call whose pool overflowed -- it falls back to the second traversal -- sizes the next call's pool
 * (forgotten again when this function is called).)
 * Bytes of workspace for the per-hit terms {W, 1/h^2} that column-density tasks record for the
 * ordered final sum; 0 (default) sizes it from the ray count (64 KiB per ray, 64 MiB to 2 GiB).
 * A pool that runs dry costs time, not correctness: the affected subtrees are walked again by the
 * launch that adds the terms up. */
int grace_b200_set_trace_pool(grace_b200_ctx* ctx, size_t bytes);
/* Hit lists (grace_b200_trace_hits_count_f4 + _fill_f4, grace_b200_trace_sorted_tiles_f4): 1 (default) = ONE
 * traversal -- the count call records the hits, the fill call copies them to the caller's arrays; 2 = two
 * traversals (count, then fill), the reference's scheme (cuda/trace_sph.cuh:112-168), also what a call falls
 * back to when the recording pool overflows.  Same lists either way. */
int grace_b200_set_hit_list_passes(grace_b200_ctx* ctx, int passes);
/* Diagnostic: counters of the last hit-count / column-density call's work stealing:
 * h_stats8 = {units finished, subtrees stolen (tasks), -, -, -, chunks of the term pool taken,
 * packets robbed (hit-list recording: slots of the copy order), hit-list recording: pool overflowed}.
 * Synchronises the stream. */
int grace_b200_trace_balance_stats(grace_b200_ctx* ctx, int* h_stats8, void* stream);
/* Device-side error flag of the last trace launch: 0 = none, 1 = traversal stack overflow
 * (the reference asserts on this only under GRACE_DEBUG, bintree_trace.cuh:162-164),
 * 2 = traversal did not terminate within 2*n_nodes steps (malformed tree).  A non-zero flag means
 * the outputs of that launch are incomplete.  The trace calls that already synchronise
 * (grace_b200_trace_hits_count_f4) check it themselves and fail with GRACE_B200_EDEVICE; the
 * asynchronous ones (hit counts, column densities, hit-list fill) cannot: poll this after them.
 * Synchronises the stream. */
int grace_b200_device_error(grace_b200_ctx* ctx, int* h_flag, void* stream);

/* All trace calls return GRACE_B200_EINVAL unless n_rays % 32 == 0
 * (bintree_trace.cuh:231-238): a packet is 32 consecutive rays. */

/* replaces: trace_hitcounts_sph (cuda/trace_sph.cuh:58-79). */
int grace_b200_trace_hitcounts_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays,
                                  size_t n_rays, const float* d_spheres4, size_t n,
                                  const grace_b200_tree* tree, int* d_hit_counts, void* stream);
/* replaces: trace_cumulative_sph (cuda/trace_sph.cuh:82-109). */
int grace_b200_trace_cumulative_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays,
                                   size_t n_rays, const float* d_spheres4, size_t n,
                                   const grace_b200_tree* tree, float* d_cumulated, void* stream);
/* replaces: trace_sph / trace_with_sentinels_sph (cuda/trace_sph.cuh:112-241), split in
 * two calls because a C ABI cannot resize the caller's vectors:
 *   _count : pass 1 (hit counts) + exclusive scan -> d_ray_offsets[n_rays];
 *            *h_total_hits = sum of counts (stream is synchronised).  With
 *            with_sentinels != 0 each offset is shifted by its ray index and the
 *            total includes one sentinel slot per ray (trace_sph.cuh:196-207).
 *   _fill  : pass 2 writes (sphere index, kernel integral, distance) per hit in
 *            emission order starting at d_ray_offsets[ray].  When it directly follows the
 *            _count call for the same rays and offsets array (no other call on the context in
 *            between) it is a copy: _count has recorded the hits in the workspace while it
 *            counted them, so the tree is walked once, not twice as in the reference
 *            (cuda/trace_sph.cuh:125-165).  Otherwise it walks the tree again.
 * Returns GRACE_B200_ERANGE if the total exceeds INT32_MAX (the reference's
 * offsets are int, trace_sph.cuh:117,137): tile the rays. */
int grace_b200_trace_hits_count_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays,
                                   size_t n_rays, const float* d_spheres4, size_t n,
                                   const grace_b200_tree* tree, int with_sentinels,
                                   int* d_ray_offsets, long long* h_total_hits, void* stream);
int grace_b200_trace_hits_fill_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays,
                                  size_t n_rays, const float* d_spheres4, size_t n,
                                  const grace_b200_tree* tree, const int* d_ray_offsets,
                                  int* d_hit_indices, float* d_hit_integrals,
                                  float* d_hit_distances, void* stream);

/* Traversal counters of the reference packet algorithm on these inputs (no reference
 * counterpart; used for the algorithmic-bytes figure):
 * h_stats4 = {inner-node visits, leaf visits, primitives staged, ray-sphere hits},
 * summed over all packets.  Synchronises the stream. */
int grace_b200_trace_stats_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                              const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                              long long* h_stats4, void* stream);

/* Diagnostic: work counters of the production packet kernel (hit-count mode):
 * h_prof4 = {inner-node steps, leaf visits, primitives in visited leaves, primitives kept
 * after the packet-bound cull}; with h_per_packet != 0 the buffer must hold
 * 4 + 4*(n_rays/32) values and also receives {cycles, node steps, leaf visits, kept} per
 * packet.  Synchronises the stream. */
int grace_b200_trace_packet_profile_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                       const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                       int* d_hit_counts, long long* h_prof4, int h_per_packet,
                                       void* stream);

/* Diagnostic: per-ray cost of the per-ray traversal (sphere tests and inner-node steps). */
int grace_b200_trace_ray_cost_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                 const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                 int* d_sphere_tests, int* d_node_steps, void* stream);

/* replaces: sort_by_distance (cuda/sort.cuh:100-131 -> sgpu::SegSortPairsFromIndices,
 *   external/sgpu/kernels/segmentedsort.cuh:732-779, + two order_by_index gathers).
 * Stable ascending sort of each ray's segment [offsets[r], offsets[r+1]) by distance;
 * indices and the 32-bit payload d_hit_data are permuted identically, in place. */
int grace_b200_sort_by_distance(grace_b200_ctx* ctx, float* d_hit_distances,
                                const int* d_ray_offsets, size_t n_rays, size_t total_hits,
                                int* d_hit_indices, void* d_hit_data, void* stream);

/* Sorted hit lists of a ray set too large for one trace_sph call, streamed in ray tiles
 * (replaces: the loop a caller of trace_sph + sort_by_distance has to write around the reference's
 * 32-bit offsets, cuda/trace_sph.cuh:112-168, cuda/sort.cuh:100-131; BASELINE config 4: a 4096^2
 * projection of 2^24 particles has ~5e10 hits).  Per tile of rays: count, scan, fill,
 * sort_by_distance, then `consume` -- all in library-owned buffers sized by hit_budget (hits per
 * tile, < 2^31; 12 bytes each, two sets) that are reused for every tile and kept by the context.
 * rays_per_tile (multiple of 32; 0 = 65536) is halved whenever a tile would exceed the budget.
 * `consume` is called once per tile, in ray order, with the tile's first ray, its ray count, the
 * exclusive offsets into the tile's lists, the number of hits and the three sorted arrays (device
 * memory, valid until it returns) and a stream: it must enqueue its work on THAT stream (not the
 * caller's) and may not call back into the library with this context; a non-zero return aborts.
 * The sort and the consumer of tile k overlap the counting traversal of tile k + 1.
 * *h_total_hits (may be NULL) receives the number of hits of all tiles.  Synchronises both streams. */
typedef int (*grace_b200_hits_tile_fn)(void* user, size_t first_ray, size_t n_rays, const int* d_ray_offsets,
                                       long long n_hits, const int* d_hit_indices, const float* d_hit_integrals,
                                       const float* d_hit_distances, void* stream);
int grace_b200_trace_sorted_tiles_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                     const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                     size_t hit_budget, size_t rays_per_tile,
                                     grace_b200_hits_tile_fn consume, void* user,
                                     long long* h_total_hits, void* stream);

/* The 51-entry kernel line-integral table (cuda/trace_sph.cuh:22-50). */
const double* grace_b200_kernel_integral_table(int* n_table);

/* ---- ray generators --------------------------------------------------------- */
/* replaces: uniform_random_rays / uniform_random_rays_single_octant
 *   (cuda/gen_rays.cuh:26-96 -> cuda/kernels/gen_rays.cuh:104-205,416-522):
 *   cuRAND XORWOW normals, double rnorm3d normalisation, 30-bit direction Morton
 *   key, stable sort by key.  octant < 0 = full sphere, else grace::Octants 0..7. */
int grace_b200_uniform_random_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays,
                                   float ox, float oy, float oz, float length,
                                   int octant, unsigned long long seed, void* stream);
/* replaces: one_to_many_rays (cuda/gen_rays.cuh:98-186 -> kernels/gen_rays.cuh:209-244,526-617).
 * d_points: n_rays records of point_stride_floats floats, xyz first.  For
 * GRACE_B200_ENDPOINT_SORT h_bot3/h_top3 give the key bounds (NULL -> computed). */
int grace_b200_one_to_many_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays,
                                float ox, float oy, float oz, const float* d_points,
                                int point_stride_floats, int sort_type,
                                const float* h_bot3, const float* h_top3, void* stream);
/* replaces: plane_parallel_random_rays (cuda/gen_rays.cuh:188-238 -> kernels :247-316,620-664). */
int grace_b200_plane_parallel_random_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays,
                                          int width, int height, const float* h_base3,
                                          const float* h_w3, const float* h_h3, float length,
                                          unsigned long long seed, void* stream);
/* replaces: orthographic_projection_rays (cuda/gen_rays.cuh:240-290 -> kernels :319-360,667-725). */
int grace_b200_orthographic_projection_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays,
                                            int resolution_x, int resolution_y,
                                            const float* h_camera_position3,
                                            const float* h_look_at3, const float* h_view_up3,
                                            float vertical_extent, float length, void* stream);
/* replaces: pinhole_camera_rays (cuda/gen_rays.cuh:292-399 -> kernels :362-395,727-787). */
int grace_b200_pinhole_camera_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays,
                                   int resolution_x, int resolution_y,
                                   const float* h_camera_position3, const float* h_look_at3,
                                   const float* h_view_up3, float fov_y, float length,
                                   void* stream);
/* HEALPix NESTED pixel centres as one-to-many rays from (ox,oy,oz): directions
 * pix2vec_nest(nside, first_pixel + i) (RayVectorGeneration/src/chealpix/chealpix.c:
 * 112-126,357-391,459-467), all of the given length. */
int grace_b200_healpix_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays,
                            long nside, long first_pixel, float ox, float oy, float oz,
                            float length, void* stream);

/* ---- utilities used by the drivers ------------------------------------------ */
/* Synthetic Gadget-shaped SPH snapshot: float4 {x,y,z,h} in [0,1)^3,
 * 30 % uniform background + 70 % in Plummer halos, h from the analytic local
 * density with N_ngb = 32, stored in Peano-Hilbert cell order (2^7 cells/side). */
int grace_b200_synth_gadget_f4(grace_b200_ctx* ctx, float* d_spheres4, size_t n,
                               unsigned int seed, void* stream);
/* Exclusive prefix sum of int32 (thrust::exclusive_scan, trace_sph.cuh:135);
 * d_total (device, 64-bit, may be NULL) receives the grand total. */
int grace_b200_exclusive_scan_i32(grace_b200_ctx* ctx, const int* d_in, int* d_out, size_t n,
                                  long long* d_total, void* stream);

/* ---- before the path: Gadget-2 (type 1) snapshots ------------ */
/* replaces: read_gadget (tests/helper/read_gadget.cuh:69-167): gas positions + smoothing
 * lengths -> float4 {x,y,z,h} records on the device.  Block offsets are derived from the
 * header; the data are read in chunks through pinned staging buffers with the host read of
 * chunk c+1 overlapping the H2D copy of chunk c, ordered on `stream` (asynchronous after the
 * last read: queue the build behind it).  *n_gas receives the gas particle count;
 * GRACE_B200_ERANGE if it exceeds `capacity` (records), GRACE_B200_EINVAL for an unreadable /
 * truncated file or one without gas particles (read_gadget.cuh:85-90 throws for that). */
int grace_b200_read_gadget_f4(grace_b200_ctx* ctx, const char* path, float* d_spheres4, size_t capacity,
                              size_t* n_gas, void* stream);
/* Header only: npart[6], mass[6] (either may be NULL), gas count. */
int grace_b200_gadget_info(const char* path, long long* npart6, double* mass6, long long* n_gas);
/* Writes float4 {x,y,z,h} host records as a Gadget-2 type-1 file the reference's reader loads
 * (driver utility: SURVEY.md 8d asks for the synthetic snapshot on disk so the unmodified
 * reference drivers can read the same bytes).  n_other particles of type 1 are appended to
 * every all-particle block; other_has_mass_block != 0 gives them header mass 0 and hence a
 * MASS block. */
int grace_b200_write_gadget_f4(const char* path, const float* h_spheres4, size_t n_gas, size_t n_other,
                               int other_has_mass_block);

/* ---- after the path: scans along sorted hit lists ------------- */
/* replaces: exclusive_segmented_scan (cuda/scan.cuh:15-38).  d_results[i] = sum of
 * d_data[segment start .. i); segment s = [offsets[s], offsets[s+1]) and the last one ends at
 * n_data.  d_data and d_results may be the same array.  Exact for integer-valued data (the
 * reference's test, tests/segmented_scan/segmented_scan.cu:97-140); general floats are summed
 * in 32-element warp-scan order, within rounding of the sequential sum. */
int grace_b200_exclusive_segmented_scan_f32(grace_b200_ctx* ctx, const int* d_segment_offsets,
                                            size_t n_segments, const float* d_data, size_t n_data,
                                            float* d_results, void* stream);
/* replaces: weighted_exclusive_segmented_scan (cuda/scan.cuh:45-58) with multiply_by_weights
 * (cuda/kernels/weights.cuh:13-59) fused: scans d_weights[d_weight_map[i]] * d_to_sum[i]. */
int grace_b200_weighted_exclusive_segmented_scan_f32(grace_b200_ctx* ctx, const float* d_to_sum,
                                                     const float* d_weights, const unsigned* d_weight_map,
                                                     const int* d_segment_offsets, size_t n_segments,
                                                     size_t n_data, float* d_sum, void* stream);
/* replaces: offsets_to_segments (cuda/sort.cuh:20-41), including its numbering when segments
 * are empty (count of distinct values among offsets[1..s]). */
int grace_b200_offsets_to_segments(grace_b200_ctx* ctx, const int* d_offsets, size_t n_offsets,
                                   int* d_segments, size_t n_data, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GRACE_B200_H */
