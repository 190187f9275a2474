// trace.cu -- packet traversal of the ALBVH with SPH kernel line integrals.
//
// Reference behaviour (GRACE): gpu::trace_kernel, cuda/kernels/bintree_trace.cuh:52-197
// (warp = packet of 32 consecutive rays sharing one stack; a child is pushed when ANY
// lane hits its box, right first then left; at a leaf every lane tests every
// primitive), AABBs_hit cuda/device/intersect.cuh:10-40 with the integer 3-input
// min/max of cuda/device/intrinsics.cuh:8-52, sphere_hit generic/intersect.h:10-55,
// functors cuda/functors/trace.cuh:163-235, entry points cuda/trace_sph.cuh:58-241.
// The reference caps the grid at 112 blocks x 8 warps (kernel_config.h:11), fetches a
// node with four dependent texture reads and re-uploads the LUT on every call.
//
// B200 design:
//   * one warp per packet, packets handed out by an atomic ticket to a persistent
//     grid sized to fill all 148 SMs (dynamic load balance: packet costs vary by
//     orders of magnitude on clustered SPH data);
//   * traversal stack is warp-uniform: top of stack in a register, the rest in a
//     per-warp shared-memory array;
//   * the 64-byte node is fetched with four 128-bit read-only loads issued together;
//   * leaf primitives are staged once per packet into shared memory (coalesced
//     128-bit loads) and broadcast from there;
//   * the 51-entry double-precision kernel table lives in shared memory for the whole
//     launch and in constant memory between launches (no per-call upload);
//   * every floating-point expression that decides a hit, or feeds the integral, is
//     written with explicit round-to-nearest intrinsics in the contraction pattern
//     nvcc gives the reference (DESIGN.md "numeric contract"), so hit sets are
//     bit-identical to the reference's CUDA build.
#include "common.cuh"

#include <math_constants.h>

#include <algorithm>
#include <cstring>

namespace {

constexpr int TR_THREADS = 128;          // 4 packets per CTA
constexpr int TR_WARPS = TR_THREADS / 32;
constexpr int TR_STACK = 96;             // reference STACK_SIZE is 64 (kernel_config.h:13)
constexpr int N_TABLE = 51;

enum { MODE_COUNT = 0, MODE_CUMULATIVE = 1, MODE_FILL = 2, MODE_STATS = 3, MODE_RAYCOST = 4,
       MODE_REC = 5 };      // MODE_REC: hit lists in one traversal (packet kernel only): hits recorded in chunk chains

// Numeric data of cuda/trace_sph.cuh:32-48: line integrals of the Gadget-2 cubic
// spline at impact parameter b/h = i/50.
__constant__ double c_kernel_table[N_TABLE] = {
    1.90986019771937, 1.90563449910964, 1.89304415940934, 1.87230928086763,
    1.84374947679902, 1.80776276033034, 1.76481079856299, 1.71540816859939,
    1.66011373131439, 1.59952322363667, 1.53426266082279, 1.46498233888091,
    1.39235130929287, 1.31705223652377, 1.23977618317103, 1.16121278415369,
    1.08201943664419, 1.00288866679720, 0.924475767210246, 0.847415371038733,
    0.772316688105931, 0.699736940377312, 0.630211918937167, 0.564194562399538,
    0.502076205853037, 0.444144023534733, 0.390518196140658, 0.341148855945766,
    0.295941946237307, 0.254782896476983, 0.217538645099225, 0.184059547649710,
    0.154181189781890, 0.127726122453554, 0.104505535066266,
    8.432088120445191E-002, 6.696547102921641E-002, 5.222604427168923E-002,
    3.988433820097490E-002, 2.971866601747601E-002, 2.150552303075515E-002,
    1.502124104014533E-002, 1.004371608622562E-002, 6.354242122978656E-003,
    3.739494884706115E-003, 1.993729589156428E-003, 9.212900163813992E-004,
    3.395908945333921E-004, 8.287326418242995E-005, 7.387919939044624E-006,
    0.000000000000000E+000
};
const double h_kernel_table[N_TABLE] = {
    1.90986019771937, 1.90563449910964, 1.89304415940934, 1.87230928086763,
    1.84374947679902, 1.80776276033034, 1.76481079856299, 1.71540816859939,
    1.66011373131439, 1.59952322363667, 1.53426266082279, 1.46498233888091,
    1.39235130929287, 1.31705223652377, 1.23977618317103, 1.16121278415369,
    1.08201943664419, 1.00288866679720, 0.924475767210246, 0.847415371038733,
    0.772316688105931, 0.699736940377312, 0.630211918937167, 0.564194562399538,
    0.502076205853037, 0.444144023534733, 0.390518196140658, 0.341148855945766,
    0.295941946237307, 0.254782896476983, 0.217538645099225, 0.184059547649710,
    0.154181189781890, 0.127726122453554, 0.104505535066266,
    8.432088120445191E-002, 6.696547102921641E-002, 5.222604427168923E-002,
    3.988433820097490E-002, 2.971866601747601E-002, 2.150552303075515E-002,
    1.502124104014533E-002, 1.004371608622562E-002, 6.354242122978656E-003,
    3.739494884706115E-003, 1.993729589156428E-003, 9.212900163813992E-004,
    3.395908945333921E-004, 8.287326418242995E-005, 7.387919939044624E-006,
    0.000000000000000E+000
};

// Two-box slab test, bit-for-bit AABBs_hit (cuda/device/intersect.cuh:16-39):
// t = (plane - o) * invd (FADD then FMUL), per-axis FMNMX, then the three-input
// min/max stages as SIGNED INTEGER compares on the float bit patterns, final >=.
__device__ __forceinline__ bool slab_hit(float bx, float tx, float by, float ty, float bz, float tz,
                                         float ox, float oy, float oz,
                                         float ix, float iy, float iz, float len)
{
    const float tbx = __fmul_rn(__fsub_rn(bx, ox), ix);
    const float ttx = __fmul_rn(__fsub_rn(tx, ox), ix);
    const float tby = __fmul_rn(__fsub_rn(by, oy), iy);
    const float tty = __fmul_rn(__fsub_rn(ty, oy), iy);
    const float tbz = __fmul_rn(__fsub_rn(bz, oz), iz);
    const float ttz = __fmul_rn(__fsub_rn(tz, oz), iz);
    const int zmin = max(min(__float_as_int(tbz), __float_as_int(ttz)), 0);
    const int zmax = min(max(__float_as_int(tbz), __float_as_int(ttz)), __float_as_int(len));
    const int tmin = max(max(__float_as_int(fminf(tbx, ttx)), __float_as_int(fminf(tby, tty))), zmin);
    const int tmax = min(min(__float_as_int(fmaxf(tbx, ttx)), __float_as_int(fmaxf(tby, tty))), zmax);
    return __int_as_float(tmax) >= __int_as_float(tmin);
}

// generic/intersect.h:16-48 in the SASS-verified contraction:
//   dot = fma(pz,rz, fma(px,rx, py*ry)); b_k = fma(-r_k, dot, p_k);
//   b2 = fma(bz,bz, fma(bx,bx, by*by)); hit iff !(b2 >= h*h) && !(dot < 0) && !(dot >= len)
__device__ __forceinline__ bool sphere_test(const float4 s, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float len,
                                            float& b2, float& dot)
{
    const float px = __fsub_rn(s.x, ox), py = __fsub_rn(s.y, oy), pz = __fsub_rn(s.z, oz);
    dot = __fmul_rn(py, dy);
    dot = __fmaf_rn(px, dx, dot);
    dot = __fmaf_rn(pz, dz, dot);
    const float bx = __fmaf_rn(-dx, dot, px);
    const float by = __fmaf_rn(-dy, dot, py);
    const float bz = __fmaf_rn(-dz, dot, pz);
    b2 = __fmul_rn(by, by);
    b2 = __fmaf_rn(bx, bx, b2);
    b2 = __fmaf_rn(bz, bz, b2);
    const float r2 = __fmul_rn(s.w, s.w);
    return !(b2 >= r2) && !(dot < 0.0f) && !(dot >= len);
}

// cuda/functors/trace.cuh:183-186 + generic/interpolate.h:15-38 (device branch).
__device__ __forceinline__ float kernel_integral(float b2, float h, const double* table)
{
    const float ir = __fdiv_rn(1.0f, h);
    float x = __fmul_rn(__fmul_rn(__fsqrt_rn(b2), ir), 50.0f);
    int i = __float2int_rz(x);
    if (i >= N_TABLE - 1) { x = (float)(N_TABLE - 1); i = N_TABLE - 2; }
    i = max(i, 0);
    const double y0 = table[i], y1 = table[i + 1];
    const double t = __dsub_rn((double)x, (double)i);
    const double y = __fma_rn(t, __dsub_rn(y1, y0), y0);
    return __fmul_rn((float)y, __fmul_rn(ir, ir));
}

// OnHit_sphere_cumulate (cuda/functors/trace.cuh:183-191): nvcc contracts
// `integral *= ir*ir; data += integral;` into one FFMA, data = fma((float)y, ir*ir, data)
// (SASS of the reference's cumulative kernel: F2F.F32.F64 ; FFMA R27, R8, R15, R27).
__device__ __forceinline__ float kernel_accumulate(float cum, float b2, float h, const double* table)
{
    const float ir = __fdiv_rn(1.0f, h);
    float x = __fmul_rn(__fmul_rn(__fsqrt_rn(b2), ir), 50.0f);
    int i = __float2int_rz(x);
    if (i >= N_TABLE - 1) { x = (float)(N_TABLE - 1); i = N_TABLE - 2; }
    i = max(i, 0);
    const double y0 = table[i], y1 = table[i + 1];
    const double t = __dsub_rn((double)x, (double)i);
    const double y = __fma_rn(t, __dsub_rn(y1, y0), y0);
    return __fmaf_rn((float)y, __fmul_rn(ir, ir), cum);
}

template <int MODE>
__global__ void __launch_bounds__(TR_THREADS)
trace_kernel(const grace_b200_ray* __restrict__ rays, int n_packets,
             const float4* __restrict__ spheres,
             const int4* __restrict__ nodes, const int4* __restrict__ leaves,
             int n_nodes, const int* __restrict__ root_ptr, int max_per_leaf,
             int* __restrict__ out_counts, float* __restrict__ out_cum,
             const int* __restrict__ offsets, int* __restrict__ hit_idx,
             float* __restrict__ hit_integral, float* __restrict__ hit_dist,
             int* __restrict__ packet_counter, int* __restrict__ err_flag,
             unsigned long long* __restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char tr_smem[];
    // layout: [double table 51 (+pad)] [stacks TR_WARPS x TR_STACK int] [prims TR_WARPS x mpl float4]
    double* s_table = (double*)tr_smem;
    int* s_stack_all = (int*)(tr_smem + 52 * sizeof(double));
    float4* s_prims_all = (float4*)(s_stack_all + TR_WARPS * TR_STACK);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE == MODE_CUMULATIVE || MODE == MODE_FILL) {
        for (int i = threadIdx.x; i < N_TABLE; i += TR_THREADS) s_table[i] = c_kernel_table[i];
        __syncthreads();
    }
    int* stack = s_stack_all + warp * TR_STACK;
    float4* s_prims = s_prims_all + warp * max_per_leaf;
    const int root = __ldg(root_ptr);

    for (;;) {
        int packet = 0;
        if (lane == 0) packet = atomicAdd(packet_counter, 1);
        packet = __shfl_sync(0xffffffffu, packet, 0);
        if (packet >= n_packets) break;
        const int ray_index = packet * 32 + lane;

        const grace_b200_ray ray = rays[ray_index];
        const float ix = __fdiv_rn(1.0f, ray.dx), iy = __fdiv_rn(1.0f, ray.dy),
                    iz = __fdiv_rn(1.0f, ray.dz);
        int count = 0;
        float cum = 0.0f;
        int cursor = 0;
        if (MODE == MODE_FILL) cursor = offsets[ray_index];

        unsigned long long st_nodes = 0, st_leaves = 0, st_prims = 0;
        int sp = 0;            // number of entries below the register top
        int top = root;        // top of stack lives in a register; -1 = empty
        while (top >= 0) {
            if (top < n_nodes) {
                if (MODE == MODE_STATS) ++st_nodes;
                const int4* np = nodes + 4 * (size_t)top;
                const int4 n0 = __ldg(np + 0);
                const int4 n1 = __ldg(np + 1);
                const int4 n2 = __ldg(np + 2);
                const int4 n3 = __ldg(np + 3);
                const bool hitL = slab_hit(__int_as_float(n1.x), __int_as_float(n1.y),
                                           __int_as_float(n1.z), __int_as_float(n1.w),
                                           __int_as_float(n3.x), __int_as_float(n3.y),
                                           ray.ox, ray.oy, ray.oz, ix, iy, iz, ray.length);
                const bool hitR = slab_hit(__int_as_float(n2.x), __int_as_float(n2.y),
                                           __int_as_float(n2.z), __int_as_float(n2.w),
                                           __int_as_float(n3.z), __int_as_float(n3.w),
                                           ray.ox, ray.oy, ray.oz, ix, iy, iz, ray.length);
                const bool anyL = __any_sync(0xffffffffu, hitL);
                const bool anyR = __any_sync(0xffffffffu, hitR);
                // pop, then push right, then left: left ends up on top
                if (anyL && anyR) {
                    if (sp >= TR_STACK) { if (lane == 0) *err_flag = 1; top = -1; sp = 0; continue; }
                    stack[sp++] = n0.y;
                    top = n0.x;
                } else if (anyL) {
                    top = n0.x;
                } else if (anyR) {
                    top = n0.y;
                } else {
                    top = sp > 0 ? stack[--sp] : -1;
                }
            } else {
                const int4 leaf = __ldg(leaves + (top - n_nodes));
                top = sp > 0 ? stack[--sp] : -1;
                if (MODE == MODE_STATS) { ++st_leaves; st_prims += leaf.y; }
                for (int i = lane; i < leaf.y; i += 32) s_prims[i] = __ldg(spheres + leaf.x + i);
                __syncwarp();
                for (int i = 0; i < leaf.y; ++i) {
                    const float4 s = s_prims[i];
                    float b2, dot;
                    if (sphere_test(s, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, ray.length,
                                    b2, dot)) {
                        if (MODE == MODE_COUNT || MODE == MODE_STATS) {
                            ++count;
                        } else if (MODE == MODE_CUMULATIVE) {
                            cum = kernel_accumulate(cum, b2, s.w, s_table);
                        } else {
                            hit_idx[cursor] = leaf.x + i;
                            hit_integral[cursor] = kernel_integral(b2, s.w, s_table);
                            hit_dist[cursor] = dot;
                            ++cursor;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (MODE == MODE_COUNT) out_counts[ray_index] = count;
        if (MODE == MODE_CUMULATIVE) out_cum[ray_index] = cum;
        if (MODE == MODE_STATS) {
            unsigned hits = __reduce_add_sync(0xffffffffu, (unsigned)count);
            if (lane == 0) {
                atomicAdd(stats + 0, st_nodes);
                atomicAdd(stats + 1, st_leaves);
                atomicAdd(stats + 2, st_prims);
                atomicAdd(stats + 3, (unsigned long long)hits);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Per-ray traversal (default).
//
// Every (ray, sphere) pair whose sphere_test() is true must be found; which OTHER pairs
// get tested is free.  The reference's packet rule (test a leaf on all 32 lanes if any
// lane's slab test reaches it) makes one beam that engulfs a dense halo test every
// particle in it on every lane -- measured on the 2^24-particle workload: 11.3k tests per
// ray for 1.3k hits, and a few such packets keep 83 % of the SMs idle at the tail.
// Here each lane walks the tree for its own ray with a slab test against boxes PADDED by
// pad = 64 * 2^-24 * (|ox|+|oy|+|oz|+len).  The padding exceeds every rounding error of
// sphere_test() and of the slab arithmetic (DESIGN.md "conservative slab test"), so a
// sphere that sphere_test() accepts always lies in a visited leaf: the hit set equals the
// brute-force set the reference's own test demands (tests/tree_traversal/tree_traversal.cu:
// 65-121), independent of how rays are grouped.  Left-first depth-first order visits
// leaves, hence primitives, in ascending index order -- the same order in which the
// reference accumulates -- so column densities and hit lists are bit-identical too.
//
// Stack: 32 entries per lane in shared memory laid out [depth][thread] (bank == lane,
// conflict-free at any mix of depths), deeper levels spill to a local array.
// ---------------------------------------------------------------------------
constexpr int RT_THREADS = 128;
constexpr int RT_SMEM_DEPTH = 32;
constexpr int RT_LOCAL_DEPTH = 64;

__device__ __forceinline__ bool slab_hit_padded(float bx, float tx, float by, float ty, float bz, float tz,
                                                float lox, float loy, float loz,   // o + pad
                                                float hix, float hiy, float hiz,   // o - pad
                                                float ix, float iy, float iz, float len)
{
    const float tbx = (bx - lox) * ix, ttx = (tx - hix) * ix;
    const float tby = (by - loy) * iy, tty = (ty - hiy) * iy;
    const float tbz = (bz - loz) * iz, ttz = (tz - hiz) * iz;
    const float tmin = fmaxf(fmaxf(fminf(tbx, ttx), fminf(tby, tty)), fmaxf(fminf(tbz, ttz), 0.0f));
    const float tmax = fminf(fminf(fmaxf(tbx, ttx), fmaxf(tby, tty)), fminf(fmaxf(tbz, ttz), len));
    return tmax >= tmin;
}

// One ray's depth-first walk from (cur, sp): padded slab test, left child first, so primitives are
// met in ascending index order.  Stack: RT_SMEM_DEPTH entries per lane in shared memory laid out
// [depth][thread], deeper levels in a local array.
struct RtRay { float lox, loy, loz, hix, hiy, hiz, ix, iy, iz; };
template <int MODE>
__device__ __forceinline__ void rt_walk(int cur, int sp, int* my_stack, int* lstack, const grace_b200_ray& ray, const RtRay& R,
                                        const float4* __restrict__ spheres, const int4* __restrict__ nodes,
                                        const int4* __restrict__ leaves, const int n_nodes, int* err_flag,
                                        const double* s_table, int* hit_idx, float* hit_integral, float* hit_dist,
                                        int& count, float& cum, int& cursor)
{
    const float lox = R.lox, loy = R.loy, loz = R.loz, hix = R.hix, hiy = R.hiy, hiz = R.hiz, ix = R.ix, iy = R.iy, iz = R.iz;
    while (cur >= 0) {
        // ---- inner nodes ----
        while ((unsigned)cur < (unsigned)n_nodes) {
            if (MODE == MODE_RAYCOST) ++cursor;      // node steps
            const int4* np = nodes + 4 * (size_t)cur;
            const int4 n0 = __ldg(np + 0);
            const int4 n1 = __ldg(np + 1);
            const int4 n2 = __ldg(np + 2);
            const int4 n3 = __ldg(np + 3);
            const bool hitL = slab_hit_padded(__int_as_float(n1.x), __int_as_float(n1.y),
                                              __int_as_float(n1.z), __int_as_float(n1.w),
                                              __int_as_float(n3.x), __int_as_float(n3.y),
                                              lox, loy, loz, hix, hiy, hiz, ix, iy, iz, ray.length);
            const bool hitR = slab_hit_padded(__int_as_float(n2.x), __int_as_float(n2.y),
                                              __int_as_float(n2.z), __int_as_float(n2.w),
                                              __int_as_float(n3.z), __int_as_float(n3.w),
                                              lox, loy, loz, hix, hiy, hiz, ix, iy, iz, ray.length);
            if (hitL) {
                if (hitR) {          // push right, descend left
                    if (sp < RT_SMEM_DEPTH) my_stack[sp * RT_THREADS] = n0.y;
                    else if (sp < RT_SMEM_DEPTH + RT_LOCAL_DEPTH) lstack[sp - RT_SMEM_DEPTH] = n0.y;
                    else { *err_flag = 1; }
                    ++sp;
                }
                cur = n0.x;
            } else if (hitR) {
                cur = n0.y;
            } else {
                if (sp > 0) {
                    --sp;
                    cur = sp < RT_SMEM_DEPTH ? my_stack[sp * RT_THREADS]
                                             : lstack[min(sp - RT_SMEM_DEPTH, RT_LOCAL_DEPTH - 1)];
                } else cur = -1;
            }
        }
        // ---- leaf ----
        if (cur >= n_nodes) {
            const int2 leaf = __ldg((const int2*)(leaves + (cur - n_nodes)));
            if (MODE == MODE_RAYCOST) count += leaf.y;   // sphere tests
            for (int i = 0; i < leaf.y && MODE != MODE_RAYCOST; ++i) {
                const float4 s = __ldg(spheres + leaf.x + i);
                float b2, dot;
                if (sphere_test(s, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, ray.length, b2, dot)) {
                    if (MODE == MODE_COUNT) {
                        ++count;
                    } else if (MODE == MODE_CUMULATIVE) {
                        cum = kernel_accumulate(cum, b2, s.w, s_table);
                    } else {
                        hit_idx[cursor] = leaf.x + i;
                        hit_integral[cursor] = kernel_integral(b2, s.w, s_table);
                        hit_dist[cursor] = dot;
                        ++cursor;
                    }
                }
            }
            if (sp > 0) {
                --sp;
                cur = sp < RT_SMEM_DEPTH ? my_stack[sp * RT_THREADS]
                                         : lstack[min(sp - RT_SMEM_DEPTH, RT_LOCAL_DEPTH - 1)];
            } else cur = -1;
        }
    }
}

// Largest |coordinate| of any box plane of the tree (the root's child boxes contain all others): part of the
// padding, see trace_packet.cuh.
__device__ __forceinline__ float rt_tree_extent(const int4* __restrict__ nodes, int root)
{
    const int4* np = nodes + 4 * (size_t)root;
    float m = 0.0f;
    for (int k = 1; k < 4; ++k) {
        const int4 q = __ldg(np + k);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(__int_as_float(q.x)), fabsf(__int_as_float(q.y))), fmaxf(fabsf(__int_as_float(q.z)), fabsf(__int_as_float(q.w)))));
    }
    return m < 3.0e38f ? m : 0.0f;
}

template <int MODE>
__global__ void __launch_bounds__(RT_THREADS)
trace_ray_kernel(const grace_b200_ray* __restrict__ rays, int n_packets,
                 const float4* __restrict__ spheres,
                 const int4* __restrict__ nodes, const int4* __restrict__ leaves,
                 int n_nodes, const int* __restrict__ root_ptr,
                 int* __restrict__ out_counts, float* __restrict__ out_cum,
                 const int* __restrict__ offsets, int* __restrict__ hit_idx,
                 float* __restrict__ hit_integral, float* __restrict__ hit_dist,
                 int* __restrict__ packet_counter, int* __restrict__ err_flag)
{
    __shared__ double s_table[52];
    __shared__ int s_stack[RT_SMEM_DEPTH * RT_THREADS];
    const int lane = threadIdx.x & 31;
    if (MODE == MODE_CUMULATIVE || MODE == MODE_FILL) {
        for (int i = threadIdx.x; i < N_TABLE; i += RT_THREADS) s_table[i] = c_kernel_table[i];
        __syncthreads();
    }
    int* my_stack = s_stack + threadIdx.x;
    int lstack[RT_LOCAL_DEPTH];
    const int root = __ldg(root_ptr);
    const float tree_extent = rt_tree_extent(nodes, root);

    for (;;) {
        int packet = 0;
        if (lane == 0) packet = atomicAdd(packet_counter, 1);
        packet = __shfl_sync(0xffffffffu, packet, 0);
        if (packet >= n_packets) break;
        const int ray_index = packet * 32 + lane;
        const grace_b200_ray ray = rays[ray_index];
        const float ix = __fdiv_rn(1.0f, ray.dx), iy = __fdiv_rn(1.0f, ray.dy),
                    iz = __fdiv_rn(1.0f, ray.dz);
        const float pad = 64.0f * 5.9604645e-8f * (fabsf(ray.ox) + fabsf(ray.oy) + fabsf(ray.oz) + fabsf(ray.length) + tree_extent);
        const float lox = ray.ox + pad, loy = ray.oy + pad, loz = ray.oz + pad;
        const float hix = ray.ox - pad, hiy = ray.oy - pad, hiz = ray.oz - pad;
        int count = 0;
        float cum = 0.0f;
        int cursor = 0;
        if (MODE == MODE_FILL) cursor = offsets[ray_index];

        RtRay R;
        R.lox = lox; R.loy = loy; R.loz = loz; R.hix = hix; R.hiy = hiy; R.hiz = hiz; R.ix = ix; R.iy = iy; R.iz = iz;
        rt_walk<MODE>(root, 0, my_stack, lstack, ray, R, spheres, nodes, leaves, n_nodes, err_flag, s_table, hit_idx,
                      hit_integral, hit_dist, count, cum, cursor);
        if (MODE == MODE_COUNT) out_counts[ray_index] = count;
        if (MODE == MODE_CUMULATIVE) out_cum[ray_index] = cum;
        if (MODE == MODE_RAYCOST) { out_counts[ray_index] = count; hit_idx[ray_index] = cursor; }
    }
}

#include "trace_packet.cuh"
#include "rec_lists.cuh"

size_t trace_smem_bytes(int max_per_leaf)
{
    return 52 * sizeof(double) + (size_t)TR_WARPS * TR_STACK * sizeof(int) +
           (size_t)TR_WARPS * max_per_leaf * sizeof(float4);
}

template <int MODE, int M4>
int launch_packet(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, int n_packets,
                  const float* d_spheres4, const grace_b200_tree* tree, int* out_counts,
                  float* out_cum, const int* offsets, int* hit_idx, float* hit_integral,
                  float* hit_dist, cudaStream_t st, unsigned long long* d_prof)
{
    if (MODE == MODE_STATS || MODE == MODE_RAYCOST) return GRACE_B200_EINVAL;   // other kernels
    constexpr int KMODE = (MODE == MODE_STATS || MODE == MODE_RAYCOST) ? MODE_COUNT : MODE;
    constexpr bool SUB = KMODE != MODE_FILL;          // work stealing inside the launch (counts, column densities, recorded hit lists)
    constexpr bool CHAIN = KMODE == MODE_CUMULATIVE;  // ordered term chains + fold launch
    constexpr bool RECM = KMODE == MODE_REC;          // hit lists in one traversal: hits recorded in chains, copied out by the fill call
    // the per-packet profile counters exist only in a separate MODE_COUNT instantiation
    auto kernel = (KMODE == MODE_COUNT && d_prof) ? trace_packet_kernel<KMODE, M4, KMODE == MODE_COUNT, false>
                                                  : trace_packet_kernel<KMODE, M4, false, false>;
    constexpr size_t psmem = packet_smem_bytes<KMODE, M4>();
    GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    int per_sm = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PK_THREADS, psmem));
    if (per_sm < 1) per_sm = 1;
    const int full_grid = ctx->sm_count * per_sm;
    int blocks = full_grid;
    const int need = (n_packets + PK_WARPS - 1) / PK_WARPS;
    if (blocks > need) blocks = need;
    int* counter = ctx->d_scalars + GB_SC_TRACE_CTR;
    // Load balancing (disabled for the profile counters).  Counts, column densities: one launch in which
    // units donate subtrees to idle warps, + the fold launch for column densities.  Hit lists: four
    // launches, 0 = packets, then suspended traversals resumed as tasks over 8, 2 and finally 1 ray(s).
    const bool split = (d_prof == nullptr) && (ctx->trace_budget & 0x3fffffff) > 0 && n_packets >= 2;
    GB_CUDA(cudaMemsetAsync(ctx->d_scalars + GB_SC_ERRFLAG, 0, sizeof(int), st));
    struct L2Window {       // node array persisting in L2 for the launches of this call
        cudaStream_t st; bool on;
        L2Window(grace_b200_ctx* c, const grace_b200_tree* t, cudaStream_t s) : st(s), on(false)
        {
            if (!c->l2_persist || !c->l2_persist_max || !c->l2_window_max) return;
            cudaStreamAttrValue a = {};
            a.accessPolicyWindow.base_ptr = const_cast<void*>(t->d_nodes);
            a.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)(t->n_leaves - 1) * 64, c->l2_window_max);
            a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2_persist_max / (double)a.accessPolicyWindow.num_bytes);
            a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            on = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess;
            if (!on) cudaGetLastError();
        }
        ~L2Window()
        {
            if (!on) return;
            cudaStreamAttrValue a = {};
            a.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
        }
    } l2_window(ctx, tree, st);
    PkArgs P;
    P.rays = d_rays; P.n_packets = n_packets; P.spheres = (const float4*)d_spheres4;
    P.nodes = (const int4*)tree->d_nodes; P.leaves = (const int4*)tree->d_leaves;
    P.n_nodes = tree->n_leaves - 1; P.root_ptr = tree->d_root;
    P.out_counts = out_counts; P.out_cum = out_cum; P.offsets = offsets;
    P.hit_idx = hit_idx; P.hit_integral = hit_integral; P.hit_dist = hit_dist;
    P.unit_counter = counter; P.err_flag = ctx->d_scalars + GB_SC_ERRFLAG; P.prof = d_prof;
    if (!split && RECM) return GRACE_B200_EINVAL;      // recording needs the pool: the caller falls back to two passes
    if (!split) {
        PkTasks T = {};
        T.kind = PK_KIND_PACKETS;
        GB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
        kernel<<<blocks, PK_THREADS, psmem, st>>>(P, T);
        GB_LAUNCH_CHECK();
        return GRACE_B200_OK;
    }
    if (SUB) {
        // work stealing inside one launch: one slot per resident warp, theft records, roots, term pool
        const int n_slots = full_grid * PK_WARPS;
        const int records_cap = 1 << 20;
        const size_t rec_bytes = gb_align((size_t)records_cap * PK_DREC_WORDS * 4);
        const size_t slot_bytes = gb_align((size_t)n_slots * sizeof(int));
        const size_t roots_bytes = (CHAIN || RECM) ? gb_align((size_t)n_packets * sizeof(int2)) : 0;
        const size_t rcum_bytes = (CHAIN || RECM) ? gb_align((size_t)n_packets * 32 * sizeof(float)) : 0;     // hit records: own counts (int)
        // term pool: the terms of stolen work only -- at most the hits of the packets in flight when the
        // tickets run out, i.e. bounded by the grid, not by the ray count.  A dry pool costs time (the
        // fold launch walks the aborted tasks itself), never correctness.
        size_t pool_bytes = 0;
        int pool_cap = 0;
        if (RECM) {
            // every hit of the call, 16 bytes each, + one partly filled chunk per unit.  Overflow = this call falls back
            // to the two-pass fill and the next one asks for what this one would have needed (rec_pool_learned).
            pool_bytes = ctx->rec_pool_hint ? ctx->rec_pool_hint : ctx->trace_pool_bytes ? ctx->trace_pool_bytes
                                            : std::min<size_t>(std::max<size_t>((size_t)n_packets * 32 * 32768, (size_t)256 << 20), (size_t)4096 << 20);
            pool_bytes = std::max(pool_bytes, ctx->rec_pool_learned);
            pool_cap = (int)std::min<size_t>(pool_bytes / PK_RCH_BYTES, 1u << 30);
            pool_bytes = gb_align((size_t)pool_cap * PK_RCH_BYTES);
        }
        if (CHAIN) {
            pool_bytes = ctx->trace_pool_bytes ? ctx->trace_pool_bytes
                                               : std::min<size_t>(std::max<size_t>((size_t)n_packets * 32 * 65536, (size_t)64 << 20),
                                                                  (size_t)2048 << 20);
            pool_cap = (int)std::min<size_t>(pool_bytes / PK_CH_BYTES, 1u << 30);
            pool_bytes = gb_align((size_t)pool_cap * PK_CH_BYTES);
        }
        const size_t adv_bytes = gb_align(PK_ADV * sizeof(int));
        const size_t order_bytes = RECM ? gb_align((size_t)pool_cap * PK_COPY_G * sizeof(int)) : 0;       // hit records: the copy order (groups of chunks)
        char* w = (char*)gb_workspace(ctx, GB_WS_HEAD + rec_bytes + 2 * slot_bytes + adv_bytes + roots_bytes + rcum_bytes + order_bytes + pool_bytes);
        if (!w) return GRACE_B200_ENOMEM;
        int* lb = ctx->d_scalars + GB_SC_LB;
        PkTasks T = {};
        T.kind = PK_KIND_PACKETS;
        char* p = w + GB_WS_HEAD;
        T.records = (int*)p; p += rec_bytes;
        T.state = (int*)p; p += slot_bytes;
        T.resp = (int*)p; p += slot_bytes;
        T.adv = (int*)p; p += adv_bytes;
        T.roots = (int2*)p; p += roots_bytes;
        T.root_cum = (float*)p; p += rcum_bytes;
        T.order = (int*)p; p += order_bytes;
        T.pool = (CHAIN || RECM) ? p : nullptr;
        T.n_slots = n_slots;
        T.records_cap = records_cap;
        // default: a unit can be robbed after 16 steps (2-4 % better than 64 from 2^12 to 2^20 rays); the hit-list rounds keep 64
        T.budget = ctx->trace_budget_set ? (ctx->trace_budget & 0x3fffffff) : 16;
        T.eager = (ctx->trace_budget & GRACE_B200_BUDGET_EAGER) ? 1 : 0;
        T.lb = lb;
        T.n_roots = lb + 6;
        T.pool_ctr = lb + 5; T.pool_cap = pool_cap;
        GB_CUDA(cudaMemsetAsync(lb, 0, 8 * sizeof(int), st));
        GB_CUDA(cudaMemsetAsync(T.state, 0, slot_bytes, st));         // nobody walking
        GB_CUDA(cudaMemsetAsync(T.adv, 0, adv_bytes, st));
        GB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
        // robbed packets and their tasks add into the same cells
        if (KMODE == MODE_COUNT || RECM) GB_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)n_packets * 32 * sizeof(int), st));
        if (RECM) GB_CUDA(cudaMemsetAsync(T.order, 0xff, order_bytes, st));      // -1: no chunk
        // the whole grid: the warps that get no packet are the first thieves
        kernel<<<full_grid, PK_THREADS, psmem, st>>>(P, T);
        GB_LAUNCH_CHECK();
        if (RECM) {      // the copy launch comes with the fill call: keep what it needs
            static_assert(sizeof(PkArgs) + sizeof(PkTasks) <= sizeof(ctx->rec_blob), "rec_blob too small");
            T.kind = PK_KIND_FOLD;
            T.state = nullptr;
            memcpy(ctx->rec_blob, &P, sizeof(PkArgs));
            memcpy(ctx->rec_blob + sizeof(PkArgs), &T, sizeof(PkTasks));
            ctx->rec_valid = 1;
            ctx->rec_epoch = ctx->ws_epoch;
        }
        if (CHAIN) {
            auto fold = trace_packet_kernel<KMODE, M4, false, CHAIN>;
            GB_CUDA(cudaFuncSetAttribute(fold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            T.kind = PK_KIND_FOLD;
            T.state = nullptr;
            GB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
            fold<<<full_grid, PK_THREADS, psmem, st>>>(P, T);
            GB_LAUNCH_CHECK();
        }
        return GRACE_B200_OK;
    }
    const int records_cap = 16384, tasks_cap = 4 * records_cap;
    const size_t rec_bytes = gb_align((size_t)records_cap * PK_REC_WORDS * 4);
    // ---- hit lists: ray-subset rounds ----
    const size_t list_bytes = gb_align((size_t)tasks_cap * 8);
    char* w = (char*)gb_workspace(ctx, GB_WS_HEAD + rec_bytes + 2 * list_bytes);
    if (!w) return GRACE_B200_ENOMEM;
    int* records = (int*)(w + GB_WS_HEAD);
    int2* lists[2] = { (int2*)(w + GB_WS_HEAD + rec_bytes), (int2*)(w + GB_WS_HEAD + rec_bytes + list_bytes) };
    int* n_counts = ctx->d_scalars + GB_SC_TASKS;      // [0] records, [1] list 0, [2] list 1, [4,5] step sum (64-bit), [6] units done
    GB_CUDA(cudaMemsetAsync(n_counts, 0, 8 * sizeof(int), st));
    const int widths[3] = { 8, 2, 1 };
    const int n_rounds = 4;
    for (int round = 0; round < n_rounds; ++round) {
        PkTasks T = {};
        T.kind = round == 0 ? PK_KIND_PACKETS : PK_KIND_TASKS;
        T.records = records; T.n_records = n_counts; T.records_cap = records_cap; T.tasks_cap = tasks_cap;
        T.budget = ctx->trace_budget & 0x3fffffff;
        T.eager = (ctx->trace_budget & GRACE_B200_BUDGET_EAGER) ? 1 : 0;
        // Round 0 whose children (32/8 per packet) would all fit on idle warps: split on the budget
        // alone.  Otherwise a unit must also be heavier than the launch's running mean.
        // (3/4 of the slots: measured at 2^24 particles, 512 packets want the budget rule -- 6.8 vs
        // 12 ms -- and 1024 packets the mean rule -- 10.5 vs 14.4 ms -- on 3552 and on 4144 warp slots)
        const bool roomy = round == 0 && 4 * (size_t)n_packets * (32 / widths[0]) <= 3 * (size_t)full_grid * PK_WARPS;
        T.sum_steps = (unsigned long long*)(n_counts + 4); T.n_done = roomy ? nullptr : n_counts + 6;
        if (round > 0) GB_CUDA(cudaMemsetAsync(n_counts + 4, 0, 3 * sizeof(int), st));        // per-round statistics
        if (round > 0) { T.tasks_in = lists[(round - 1) & 1]; T.n_tasks_in = n_counts + 1 + ((round - 1) & 1); }
        if (round < n_rounds - 1) {
            T.tasks_out = lists[round & 1]; T.n_tasks_out = n_counts + 1 + (round & 1);
            T.child_width = widths[round];
            if (round >= 2) GB_CUDA(cudaMemsetAsync(T.n_tasks_out, 0, sizeof(int), st));   // list reuse
        }
        GB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
        kernel<<<round == 0 ? blocks : full_grid, PK_THREADS, psmem, st>>>(P, T);
        GB_LAUNCH_CHECK();
    }
    return GRACE_B200_OK;
}

template <int MODE>
int launch_trace(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                 const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                 int* out_counts, float* out_cum, const int* offsets, int* hit_idx,
                 float* hit_integral, float* hit_dist, cudaStream_t st,
                 unsigned long long* d_stats = nullptr)
{
    GB_REQUIRE(ctx && (d_rays || n_rays == 0) && d_spheres4 && tree && tree->d_nodes && tree->d_leaves && tree->d_root,
               GRACE_B200_EINVAL, "NULL argument");
    // bintree_trace.cuh:231-238
    GB_REQUIRE(n_rays % 32 == 0, GRACE_B200_EINVAL,
               "Number of rays must be a multiple of the warp size (32).");
    GB_REQUIRE(n_rays < (1ull << 31), GRACE_B200_ERANGE, "more than 2^31 rays in one call");
    GB_REQUIRE(tree->n_leaves >= 2 && tree->max_per_leaf >= 1, GRACE_B200_EINVAL, "malformed tree");
    (void)n;
    if (n_rays == 0) return GRACE_B200_OK;
    const int n_packets = (int)(n_rays / 32);
    int* counter = ctx->d_scalars + GB_SC_TRACE_CTR;
    if (MODE == MODE_RAYCOST || (MODE != MODE_STATS && ctx->trace_mode == GRACE_B200_TRACE_PER_RAY)) {
        int per_sm = 0;
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_ray_kernel<MODE>, RT_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        int blocks = ctx->sm_count * per_sm;
        const int need = (n_packets + RT_THREADS / 32 - 1) / (RT_THREADS / 32);
        if (blocks > need) blocks = need;
        GB_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));
        trace_ray_kernel<MODE><<<blocks, RT_THREADS, 0, st>>>(
            d_rays, n_packets, (const float4*)d_spheres4, (const int4*)tree->d_nodes,
            (const int4*)tree->d_leaves, tree->n_leaves - 1, tree->d_root,
            out_counts, out_cum, offsets, hit_idx, hit_integral, hit_dist, counter,
            ctx->d_scalars + GB_SC_ERRFLAG);
        GB_LAUNCH_CHECK();
        return GRACE_B200_OK;
    }
    if (MODE != MODE_STATS && MODE != MODE_RAYCOST &&
        ctx->trace_mode == GRACE_B200_TRACE_PACKET && tree->max_per_leaf <= 128) {
        if (tree->max_per_leaf <= 32)
            return launch_packet<MODE, 32>(ctx, d_rays, n_packets, d_spheres4, tree, out_counts, out_cum,
                                           offsets, hit_idx, hit_integral, hit_dist, st, d_stats);
        if (tree->max_per_leaf <= 64)
            return launch_packet<MODE, 64>(ctx, d_rays, n_packets, d_spheres4, tree, out_counts, out_cum,
                                           offsets, hit_idx, hit_integral, hit_dist, st, d_stats);
        return launch_packet<MODE, 128>(ctx, d_rays, n_packets, d_spheres4, tree, out_counts, out_cum,
                                        offsets, hit_idx, hit_integral, hit_dist, st, d_stats);
    }
    const size_t smem = trace_smem_bytes(tree->max_per_leaf);
    GB_REQUIRE(smem <= 200 * 1024, GRACE_B200_EINVAL, "max_per_leaf too large for shared memory staging");
    GB_CUDA(cudaFuncSetAttribute(trace_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel<MODE>, TR_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = ctx->sm_count * per_sm;
    const int need = (n_packets + TR_WARPS - 1) / TR_WARPS;
    if (blocks > need) blocks = need;
    GB_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));   // counter + error flag
    trace_kernel<MODE><<<blocks, TR_THREADS, smem, st>>>(
        d_rays, n_packets, (const float4*)d_spheres4, (const int4*)tree->d_nodes,
        (const int4*)tree->d_leaves, tree->n_leaves - 1, tree->d_root, tree->max_per_leaf,
        out_counts, out_cum, offsets, hit_idx, hit_integral, hit_dist, counter,
        ctx->d_scalars + GB_SC_ERRFLAG, d_stats);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // namespace

// One-pass hit lists, first half: the counting traversal also records every hit {integral, distance, index} in
// chunk chains in the workspace (work stealing as for hit counts).  d_counts receives the per-ray counts.
int gb_trace_record_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays, const float* d_spheres4, size_t n,
                       const grace_b200_tree* tree, int* d_counts, cudaStream_t st)
{
    ctx->rec_valid = 0;
    if (ctx->trace_mode != GRACE_B200_TRACE_PACKET || !ctx->one_pass_lists || tree->max_per_leaf > 128 || n_rays % 32 || n_rays == 0 ||
        n_rays >= (1ull << 31))
        return GRACE_B200_EINVAL;
    (void)n;
    const int n_packets = (int)(n_rays / 32);
    int rc;
    if (tree->max_per_leaf <= 32)
        rc = launch_packet<MODE_REC, 32>(ctx, d_rays, n_packets, d_spheres4, tree, d_counts, nullptr, nullptr, nullptr, nullptr, nullptr, st, nullptr);
    else if (tree->max_per_leaf <= 64)
        rc = launch_packet<MODE_REC, 64>(ctx, d_rays, n_packets, d_spheres4, tree, d_counts, nullptr, nullptr, nullptr, nullptr, nullptr, st, nullptr);
    else
        rc = launch_packet<MODE_REC, 128>(ctx, d_rays, n_packets, d_spheres4, tree, d_counts, nullptr, nullptr, nullptr, nullptr, nullptr, st, nullptr);
    if (rc) { ctx->rec_valid = 0; return rc; }
    ctx->rec_rays = d_rays; ctx->rec_n_rays = n_rays; ctx->rec_offsets = d_counts;
    ctx->rec_valid = tree->max_per_leaf <= 32 ? 32 : tree->max_per_leaf <= 64 ? 64 : 128;
    return GRACE_B200_OK;
}

// Between the two halves, once the scan has turned the counts into offsets: the position of every stolen subtree's
// first hit, per ray (rec_resolve_kernel).  A malformed theft tree raises the overflow flag = two passes.
int gb_trace_resolve_recorded(grace_b200_ctx* ctx, const int* d_offsets, cudaStream_t st)
{
    if (!ctx->rec_valid) return GRACE_B200_OK;
    PkArgs P; PkTasks T;
    memcpy(&P, ctx->rec_blob, sizeof(PkArgs));
    memcpy(&T, ctx->rec_blob + sizeof(PkArgs), sizeof(PkTasks));
    const int blocks = std::min((P.n_packets + 3) / 4, ctx->sm_count * 16);
    rec_resolve_kernel<<<blocks, 128, 0, st>>>(T.roots, (const int*)T.root_cum, T.records, d_offsets, P.n_packets,
                                               T.lb + PK_LB_CREATED, T.records_cap, T.lb + PK_LB_OVERFLOW);
    GB_LAUNCH_CHECK();
    rec_order_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(T.pool, T.pool_ctr, T.pool_cap, T.roots, T.records, T.order);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

// Second half: copy the recorded hits to the caller's arrays, each ray's from its offset on, in emission order.
int gb_trace_copy_recorded(grace_b200_ctx* ctx, const int* d_offsets, int* d_idx, float* d_integ, float* d_dist, cudaStream_t st)
{
    if (!ctx->rec_valid) return GRACE_B200_EINVAL;
    ctx->rec_valid = 0;
    PkArgs P; PkTasks T;
    memcpy(&P, ctx->rec_blob, sizeof(PkArgs));
    memcpy(&T, ctx->rec_blob + sizeof(PkArgs), sizeof(PkTasks));
    // (per device, so on every call like the other launches: a process may drive several devices)
    GB_CUDA(cudaFuncSetAttribute(rec_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RC_WARPS * sizeof(RcWarp))));
    // groups <= chunks; the group count is in the slot counter the units drew from
    rec_copy_kernel<<<ctx->sm_count * 3, RC_WARPS * 32, RC_WARPS * sizeof(RcWarp), st>>>(T.pool, T.n_roots, T.pool_cap, T.order, T.records, d_offsets,
                                                                                        d_idx, d_integ, d_dist);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

namespace {
} // namespace

extern "C" {

int grace_b200_device_error(grace_b200_ctx* ctx, int* h_flag, void* stream)
{
    GB_REQUIRE(ctx && h_flag, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_ERRFLAG, ctx->d_scalars + GB_SC_ERRFLAG, sizeof(int),
                            cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    *h_flag = ctx->h_pinned[GB_SC_ERRFLAG];
    return GRACE_B200_OK;
}

int grace_b200_set_trace_budget(grace_b200_ctx* ctx, int steps)
{
    GB_REQUIRE(ctx && steps >= 0, GRACE_B200_EINVAL, "bad argument");
    ctx->trace_budget = steps;
    ctx->trace_budget_set = 1;
    return GRACE_B200_OK;
}

#ifdef PK_DEBUG_LB
extern "C" __attribute__((visibility("default"))) int grace_b200_debug_lb(unsigned long long* out32, int reset)
{
    if (reset) {
        unsigned long long init[32] = {};
        init[4] = init[7] = ~0ull;
        return (int)cudaMemcpyToSymbol(pk_dbg, init, sizeof(init));
    }
    return (int)cudaMemcpyFromSymbol(out32, pk_dbg, 32 * sizeof(unsigned long long));
}
#endif

int grace_b200_trace_balance_stats(grace_b200_ctx* ctx, int* h_stats8, void* stream)
{
    GB_REQUIRE(ctx && h_stats8, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_LB, ctx->d_scalars + GB_SC_LB, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) h_stats8[i] = ctx->h_pinned[GB_SC_LB + i];
    return GRACE_B200_OK;
}

int grace_b200_set_trace_pool(grace_b200_ctx* ctx, size_t bytes)
{
    GB_REQUIRE(ctx, GRACE_B200_EINVAL, "ctx is NULL");
    ctx->trace_pool_bytes = bytes;
    ctx->rec_pool_learned = 0;        // what earlier hit-list recordings would have needed is forgotten
    return GRACE_B200_OK;
}

int grace_b200_set_hit_list_passes(grace_b200_ctx* ctx, int passes)
{
    GB_REQUIRE(ctx && (passes == 1 || passes == 2), GRACE_B200_EINVAL, "passes must be 1 or 2");
    ctx->one_pass_lists = passes == 1;
    ctx->rec_valid = 0;
    return GRACE_B200_OK;
}

int grace_b200_set_trace_mode(grace_b200_ctx* ctx, int mode)
{
    GB_REQUIRE(ctx, GRACE_B200_EINVAL, "ctx is NULL");
    GB_REQUIRE(mode == GRACE_B200_TRACE_PER_RAY || mode == GRACE_B200_TRACE_PACKET || mode == GRACE_B200_TRACE_PACKET_REF, GRACE_B200_EINVAL,
               "unknown trace mode %d", mode);
    ctx->trace_mode = mode;
    return GRACE_B200_OK;
}

const double* grace_b200_kernel_integral_table(int* n_table)
{
    if (n_table) *n_table = N_TABLE;
    return h_kernel_table;
}

int grace_b200_trace_hitcounts_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                  const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                  int* d_hit_counts, void* stream)
{
    GB_REQUIRE(d_hit_counts || n_rays == 0, GRACE_B200_EINVAL, "NULL output");
    return launch_trace<MODE_COUNT>(ctx, d_rays, n_rays, d_spheres4, n, tree, d_hit_counts, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int grace_b200_trace_cumulative_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                   const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                   float* d_cumulated, void* stream)
{
    GB_REQUIRE(d_cumulated || n_rays == 0, GRACE_B200_EINVAL, "NULL output");
    return launch_trace<MODE_CUMULATIVE>(ctx, d_rays, n_rays, d_spheres4, n, tree, nullptr,
                                         d_cumulated, nullptr, nullptr, nullptr, nullptr,
                                         (cudaStream_t)stream);
}

int grace_b200_trace_stats_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                              const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                              long long* h_stats4, void* stream)
{
    GB_REQUIRE(ctx && h_stats4, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_stats = (unsigned long long*)gb_workspace(ctx, 256);
    if (!d_stats) return GRACE_B200_ENOMEM;
    GB_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), st));
    int rc = launch_trace<MODE_STATS>(ctx, d_rays, n_rays, d_spheres4, n, tree, nullptr, nullptr,
                                      nullptr, nullptr, nullptr, nullptr, st, d_stats);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(h_stats4, d_stats, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GRACE_B200_OK;
}

int grace_b200_trace_packet_profile_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                       const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                       int* d_hit_counts, long long* h_prof4, int h_per_packet,
                                       void* stream)
{
    GB_REQUIRE(ctx && h_prof4 && d_hit_counts, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t prof_words = 4 + 4 * (n_rays / 32);
    unsigned long long* d_prof = (unsigned long long*)gb_workspace(ctx, prof_words * 8 + 256);
    if (!d_prof) return GRACE_B200_ENOMEM;
    GB_CUDA(cudaMemsetAsync(d_prof, 0, prof_words * 8, st));
    const int saved = ctx->trace_mode;
    ctx->trace_mode = GRACE_B200_TRACE_PACKET;
    int rc = launch_trace<MODE_COUNT>(ctx, d_rays, n_rays, d_spheres4, n, tree, d_hit_counts, nullptr,
                                      nullptr, nullptr, nullptr, nullptr, st, d_prof);
    ctx->trace_mode = saved;
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(h_prof4, d_prof, (h_per_packet ? prof_words : 4) * sizeof(long long),
                            cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GRACE_B200_OK;
}

int grace_b200_trace_ray_cost_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                 const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                 int* d_sphere_tests, int* d_node_steps, void* stream)
{
    GB_REQUIRE(d_sphere_tests && d_node_steps, GRACE_B200_EINVAL, "NULL output");
    return launch_trace<MODE_RAYCOST>(ctx, d_rays, n_rays, d_spheres4, n, tree, d_sphere_tests, nullptr,
                                      nullptr, d_node_steps, nullptr, nullptr, (cudaStream_t)stream);
}

int grace_b200_trace_hits_fill_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                  const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                  const int* d_ray_offsets, int* d_hit_indices,
                                  float* d_hit_integrals, float* d_hit_distances, void* stream)
{
    GB_REQUIRE(d_ray_offsets && d_hit_indices && d_hit_integrals && d_hit_distances,
               GRACE_B200_EINVAL, "NULL argument");
    // the count call that produced these offsets has recorded the hits themselves (one traversal instead of two),
    // unless something else has used the workspace since or its hit pool ran dry
    if (ctx && ctx->rec_valid && ctx->rec_rays == d_rays && ctx->rec_n_rays == n_rays && ctx->rec_offsets == d_ray_offsets &&
        ctx->rec_epoch == ctx->ws_epoch)
        return gb_trace_copy_recorded(ctx, d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances, (cudaStream_t)stream);
    if (ctx) ctx->rec_valid = 0;
    return launch_trace<MODE_FILL>(ctx, d_rays, n_rays, d_spheres4, n, tree, nullptr, nullptr,
                                   d_ray_offsets, d_hit_indices, d_hit_integrals, d_hit_distances,
                                   (cudaStream_t)stream);
}

} // extern "C"
