/*
 * grace_oracle.c -- CPU restatement of GRACE's ray-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (grace-devel_b200/,
 * include/) may call, link or load this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function cites the reference file:line (paths relative to the GRACE
 * source tree, include/grace/...) whose arithmetic it restates.  Floating-point
 * expressions use fmaf() in exactly the contraction pattern nvcc 12.9 emits for
 * the reference's device code (verified with cuobjdump -sass, see DESIGN.md
 * "numeric contract"); compile with -ffp-contract=off so gcc adds none of its own.
 *
 * Parity status: pinned by (i) the reference's two Morton known-answer tests
 * (tests/morton_key/30bit_key.cu:20-26, 63bit_key.cu:20-26), (ii) the reference's
 * relational tests restated in tests/ (brute force == tree trace,
 * tests/tree_traversal/tree_traversal.cu:65-121; volume integral == 1 +- 5e-4,
 * tests/integrate/integrate.cu:53,101; sortedness, tests/distance_sort/distance_sort.cu:22-79),
 * and (iii) outputs of the reference's own CUDA build run on a B200
 * (oracle/ref_driver.cu, fixtures under tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* helpers                                                                    */
/* ------------------------------------------------------------------------- */

static inline int32_t f2i_bits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f_bits(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

/* cvt.rzi.u32.f32 / cvt.rzi.u64.f32: truncate, saturate, NaN -> 0.
 * (static_cast<KeyType>(float) in cuda/kernels/morton.cuh:46-48 compiles to
 * F2I.U32.TRUNC / F2I.U64.TRUNC.) */
static inline uint32_t cvt_rzi_u32(float f)
{
    if (!(f == f)) return 0u;
    if (f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
static inline uint64_t cvt_rzi_u64(float f)
{
    if (!(f == f)) return 0ull;
    if (f <= 0.0f) return 0ull;
    if (f >= 18446744073709551616.0f) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)f;
}

/* FMNMX semantics: NaN-suppressing, -0 < +0. */
static inline float fmin_dev(float a, float b)
{
    if (!(a == a)) return b;
    if (!(b == b)) return a;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float fmax_dev(float a, float b)
{
    if (!(a == a)) return b;
    if (!(b == b)) return a;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}

/* ------------------------------------------------------------------------- */
/* Morton keys: generic/bits.h:24-46, generic/morton.h:14-29                  */
/* ------------------------------------------------------------------------- */

ORC_API uint32_t orc_space_by_two_10bit(uint32_t x)
{
    x &= (1u << 10) - 1;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x <<  8)) & 0x0300F00Fu;
    x = (x | (x <<  4)) & 0x030C30C3u;
    x = (x | (x <<  2)) & 0x09249249u;
    return x;
}

ORC_API uint64_t orc_space_by_two_21bit(uint64_t x)
{
    x &= (1u << 21) - 1;
    x = (x | x << 32) & 0x001f00000000ffffull;
    x = (x | x << 16) & 0x001f0000ff0000ffull;
    x = (x | x <<  8) & 0x100f00f00f00f00full;
    x = (x | x <<  4) & 0x10c30c30c30c30c3ull;
    x = (x | x <<  2) & 0x1249249249249249ull;
    return x;
}

ORC_API uint32_t orc_morton_key30(uint32_t x, uint32_t y, uint32_t z)
{
    return orc_space_by_two_10bit(z) << 2 | orc_space_by_two_10bit(y) << 1
           | orc_space_by_two_10bit(x);
}

ORC_API uint64_t orc_morton_key63(uint64_t x, uint64_t y, uint64_t z)
{
    return orc_space_by_two_21bit(z) << 2 | orc_space_by_two_21bit(y) << 1
           | orc_space_by_two_21bit(x);
}

/* cuda/kernels/aabb.cuh:14-32 + cuda/util/extrema.cuh:502-513,667-678:
 * component-wise min/max of the sphere centres (CentroidSphere,
 * generic/functors/centroid.h:33-40). */
ORC_API void orc_bounds(const float* s4, long n, float* mins, float* maxs)
{
    for (int k = 0; k < 3; ++k) { mins[k] = s4[k]; maxs[k] = s4[k]; }
    for (long i = 1; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = s4[4 * i + k];
            if (v < mins[k]) mins[k] = v;
            if (v > maxs[k]) maxs[k] = v;
        }
}

/* cuda/kernels/morton.cuh:107-113 (scale, float division on the host) and
 * :46-48 (FADD, FMUL, F2I.TRUNC; no fusion possible). */
ORC_API void orc_morton_keys30(const float* s4, long n, const float* bot,
                               const float* top, uint32_t* keys)
{
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = 1023.0f / (top[k] - bot[k]);
    for (long i = 0; i < n; ++i) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k)
            q[k] = cvt_rzi_u32(scale[k] * (s4[4 * i + k] - bot[k]));
        keys[i] = orc_morton_key30(q[0], q[1], q[2]);
    }
}

ORC_API void orc_morton_keys63(const float* s4, long n, const float* bot,
                               const float* top, uint64_t* keys)
{
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = 2097151.0f / (top[k] - bot[k]);
    for (long i = 0; i < n; ++i) {
        uint64_t q[3];
        for (int k = 0; k < 3; ++k)
            q[k] = cvt_rzi_u64(scale[k] * (s4[4 * i + k] - bot[k]));
        keys[i] = orc_morton_key63(q[0], q[1], q[2]);
    }
}

/* ------------------------------------------------------------------------- */
/* Stable key sort: cuda/build_sph.cuh:46,57,70,81 (thrust::sort_by_key is a  */
/* stable LSD radix sort).  Returns the permutation: sorted[i] = in[perm[i]]. */
/* ------------------------------------------------------------------------- */

ORC_API void orc_sort_perm_u64(const uint64_t* keys, long n, int32_t* perm)
{
    if (n <= 0) return;
    int32_t* a = perm;
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) a[i] = (int32_t)i;
    for (int pass = 0; pass < 8; ++pass) {
        long count[257];
        memset(count, 0, sizeof(count));
        int shift = 8 * pass;
        for (long i = 0; i < n; ++i) count[((keys[a[i]] >> shift) & 0xFF) + 1]++;
        if (count[((keys[0] >> shift) & 0xFF) + 1] == n) continue;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (long i = 0; i < n; ++i) b[count[(keys[a[i]] >> shift) & 0xFF]++] = a[i];
        int32_t* t = a; a = b; b = t;
    }
    if (a != perm) { memcpy(perm, a, sizeof(int32_t) * (size_t)n); free(a); }
    else free(b);
}

ORC_API void orc_sort_perm_u32(const uint32_t* keys, long n, int32_t* perm)
{
    uint64_t* k64 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) k64[i] = keys[i];
    orc_sort_perm_u64(k64, n, perm);
    free(k64);
}

/* ------------------------------------------------------------------------- */
/* Deltas: cuda/kernels/albvh.cuh:33-47 (index shifted by one, N+1 outputs),  */
/* functors generic/functors/albvh.h:17-126.                                  */
/* ------------------------------------------------------------------------- */

/* DeltaEuclidean, generic/functors/albvh.h:78-80.  SASS: d = dy*dy;
 * d = fma(dx,dx,d); d = fma(dz,dz,d), differences on the raw .x/.y/.z. */
ORC_API void orc_deltas_euclid(const float* s4, long n, float* deltas)
{
    deltas[0] = INFINITY;
    deltas[n] = INFINITY;
    for (long i = 0; i + 1 < n; ++i) {
        float dx = s4[4 * i + 0] - s4[4 * (i + 1) + 0];
        float dy = s4[4 * i + 1] - s4[4 * (i + 1) + 1];
        float dz = s4[4 * i + 2] - s4[4 * (i + 1) + 2];
        float d = dy * dy;
        d = fmaf(dx, dx, d);
        d = fmaf(dz, dz, d);
        deltas[i + 1] = d;
    }
}

/* DeltaSurfaceArea, generic/functors/albvh.h:101-121 with AABBSphere
 * (generic/functors/aabb.h:9-26).  SASS: SA = Lx*Lz; SA = fma(Lx,Ly,SA);
 * SA = fma(Ly,Lz,SA). */
ORC_API void orc_deltas_sarea(const float* s4, long n, float* deltas)
{
    deltas[0] = INFINITY;
    deltas[n] = INFINITY;
    for (long i = 0; i + 1 < n; ++i) {
        const float* a = s4 + 4 * i;
        const float* b = s4 + 4 * (i + 1);
        float L[3];
        for (int k = 0; k < 3; ++k) {
            float bi = a[k] - a[3], ti = a[k] + a[3];
            float bj = b[k] - b[3], tj = b[k] + b[3];
            L[k] = fmax_dev(ti, tj) - fmin_dev(bi, bj);
        }
        float sa = L[0] * L[2];
        sa = fmaf(L[0], L[1], sa);
        sa = fmaf(L[1], L[2], sa);
        deltas[i + 1] = sa;
    }
}

/* DeltaXOR, generic/functors/albvh.h:17-47. */
ORC_API void orc_deltas_xor32(const uint32_t* keys, long n, uint32_t* deltas)
{
    deltas[0] = 0xFFFFFFFFu;
    deltas[n] = 0xFFFFFFFFu;
    for (long i = 0; i + 1 < n; ++i) deltas[i + 1] = keys[i] ^ keys[i + 1];
}
ORC_API void orc_deltas_xor64(const uint64_t* keys, long n, uint64_t* deltas)
{
    deltas[0] = ~0ull;
    deltas[n] = ~0ull;
    for (long i = 0; i + 1 < n; ++i) deltas[i + 1] = keys[i] ^ keys[i + 1];
}

/* ------------------------------------------------------------------------- */
/* ALBVH build, instantiated for float / u32 / u64 deltas.                    */
/* ------------------------------------------------------------------------- */

#define DELTA_T float
#define SUFFIX(name) name##_f32
#include "albvh_oracle.inc"
#undef DELTA_T
#undef SUFFIX

#define DELTA_T uint32_t
#define SUFFIX(name) name##_u32
#include "albvh_oracle.inc"
#undef DELTA_T
#undef SUFFIX

#define DELTA_T uint64_t
#define SUFFIX(name) name##_u64
#include "albvh_oracle.inc"
#undef DELTA_T
#undef SUFFIX

/* ------------------------------------------------------------------------- */
/* Intersection tests.                                                        */
/* ------------------------------------------------------------------------- */

typedef struct { float dx, dy, dz, ox, oy, oz, length; } orc_ray; /* ray.h:5-10 */

/* generic/intersect.h:10-55 in the FMA form nvcc emits:
 *   dot = fma(pz,rz, fma(px,rx, py*ry)); b_k = fma(-r_k, dot, p_k);
 *   b2 = fma(bz,bz, fma(bx,bx, by*by)); r2 = h*h. */
static inline int sphere_hit(const orc_ray* r, const float* s, float* b2o, float* doto)
{
    float px = s[0] - r->ox;
    float py = s[1] - r->oy;
    float pz = s[2] - r->oz;
    float dot = py * r->dy;
    dot = fmaf(px, r->dx, dot);
    dot = fmaf(pz, r->dz, dot);
    float bx = fmaf(-r->dx, dot, px);
    float by = fmaf(-r->dy, dot, py);
    float bz = fmaf(-r->dz, dot, pz);
    float b2 = by * by;
    b2 = fmaf(bx, bx, b2);
    b2 = fmaf(bz, bz, b2);
    *b2o = b2;
    *doto = dot;
    if (b2 >= s[3] * s[3]) return 0;
    if (dot < 0.0f) return 0;
    if (dot >= r->length) return 0;
    return 1;
}

ORC_API int orc_sphere_hit(const float* ray7, const float* s4, float* b2, float* dot)
{
    return sphere_hit((const orc_ray*)ray7, s4, b2, dot);
}

/* cuda/device/intrinsics.cuh:8-52: 3-input min/max on the float bit patterns
 * as signed integers. */
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline float maxf_vmaxf(float a, float b, float c)
{ return i2f_bits(imax(imax(f2i_bits(a), f2i_bits(b)), f2i_bits(c))); }
static inline float minf_vminf(float a, float b, float c)
{ return i2f_bits(imin(imin(f2i_bits(a), f2i_bits(b)), f2i_bits(c))); }
static inline float maxf_vminf(float a, float b, float c)
{ return i2f_bits(imax(imin(f2i_bits(a), f2i_bits(b)), f2i_bits(c))); }
static inline float minf_vmaxf(float a, float b, float c)
{ return i2f_bits(imin(imax(f2i_bits(a), f2i_bits(b)), f2i_bits(c))); }

/* cuda/device/intersect.cuh:10-40.  box = {bx,tx,by,ty,bz,tz}. */
static inline int box_hit(const float* invd, const float* o, float len, const float* box)
{
    float bx = (box[0] - o[0]) * invd[0];
    float tx = (box[1] - o[0]) * invd[0];
    float by = (box[2] - o[1]) * invd[1];
    float ty = (box[3] - o[1]) * invd[1];
    float bz = (box[4] - o[2]) * invd[2];
    float tz = (box[5] - o[2]) * invd[2];
    float tmin = maxf_vmaxf(fmin_dev(bx, tx), fmin_dev(by, ty), maxf_vminf(bz, tz, 0.0f));
    float tmax = minf_vminf(fmax_dev(bx, tx), fmax_dev(by, ty), minf_vmaxf(bz, tz, len));
    return tmax >= tmin;
}

/* Returns hitR + 2*hitL like AABBs_hit(); node = 16 ints in Tree layout
 * (cuda/nodes.h:21-36). */
static inline int aabbs_hit(const float* invd, const float* o, float len, const int32_t* node)
{
    const float* f = (const float*)node;
    float L[6] = { f[4], f[5], f[6], f[7], f[12], f[13] };
    float R[6] = { f[8], f[9], f[10], f[11], f[14], f[15] };
    return box_hit(invd, o, len, R) + 2 * box_hit(invd, o, len, L);
}

ORC_API int orc_aabbs_hit(const float* ray7, const int32_t* node16)
{
    const orc_ray* r = (const orc_ray*)ray7;
    float invd[3] = { 1.0f / r->dx, 1.0f / r->dy, 1.0f / r->dz };
    float o[3] = { r->ox, r->oy, r->oz };
    return aabbs_hit(invd, o, r->length, node16);
}

/* Kernel line-integral table: the numeric data of cuda/trace_sph.cuh:32-48
 * (51 doubles, impact parameter b/h = i/50). */
static const double kernel_table[51] = {
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

ORC_API const double* orc_kernel_table(void) { return kernel_table; }

/* cuda/functors/trace.cuh:183-186 + generic/interpolate.h:11-39 (device form):
 *   ir = 1/h; x = (sqrt(b2)*ir)*50; i = trunc(x) clamped; t = (double)x - i;
 *   y = fma(t, T[i+1]-T[i], T[i]) in double; integral = (float)y * (ir*ir). */
static inline float kernel_lerp(float b2, float h, float* ir2_out)
{
    float ir = 1.0f / h;
    float x = (sqrtf(b2) * ir) * 50.0f;
    int i = (int)x;                 /* F2I.TRUNC.NTZ */
    if (!(x == x)) i = 0;           /* NaN -> 0 on the device */
    if (i >= 50) { x = 50.0f; i = 49; }
    if (i < 0) i = 0;               /* unreachable for finite b2 >= 0 */
    double y0 = kernel_table[i], y1 = kernel_table[i + 1];
    double t = (double)x - (double)i;
    double y = fma(t, y1 - y0, y0);
    *ir2_out = ir * ir;
    return (float)y;
}

/* Per-hit integral as stored by OnHit_sphere_individual (cuda/functors/trace.cuh:221-228):
 * integral = (float)y * (ir*ir), one FMUL. */
static inline float kernel_integral(float b2, float h)
{
    float ir2;
    float y = kernel_lerp(b2, h, &ir2);
    return y * ir2;
}

/* OnHit_sphere_cumulate (cuda/functors/trace.cuh:183-191): `integral *= ir*ir;
 * ray_data.data += integral;` is contracted by nvcc into ONE fused multiply-add,
 * data = fma((float)y, ir*ir, data) (SASS of the reference's cumulative trace kernel:
 * F2F.F32.F64 ; FFMA R27, R8, R15, R27). */
static inline float kernel_accumulate(float cum, float b2, float h)
{
    float ir2;
    float y = kernel_lerp(b2, h, &ir2);
    return fmaf(y, ir2, cum);
}

ORC_API float orc_kernel_integral(float b2, float h) { return kernel_integral(b2, h); }

/* ------------------------------------------------------------------------- */
/* Packet traversal: cuda/kernels/bintree_trace.cuh:119-193.                  */
/* A packet is 32 consecutive rays sharing one stack; a node is pushed if ANY */
/* lane hits its box (right first, then left, so left is processed first);    */
/* at a leaf EVERY lane tests EVERY primitive.                                */
/*                                                                            */
/* mode 0: hit counts   (Intersect_sphere_bool + OnHit_increment, trace_sph.cuh:58-79)   */
/* mode 1: cumulative   (OnHit_sphere_cumulate, trace_sph.cuh:82-109)                    */
/* mode 2: per-hit fill (OnHit_sphere_individual, trace_sph.cuh:150-167); offsets = per-ray */
/*         write cursor (RayEntry_from_array)                                            */
/* stats (optional, 3 int64 per packet): node visits, leaf visits, prims staged.         */
/* Returns the maximum stack depth reached (reference STACK_SIZE is 64,                  */
/* cuda/kernel_config.h:13).                                                             */
/* ------------------------------------------------------------------------- */
ORC_API int orc_trace(const float* rays7, long n_rays, const float* s4,
                      const int32_t* nodes, const int32_t* leaves, long n_leaves,
                      int root, int mode,
                      int32_t* out_counts, float* out_cum,
                      const int32_t* offsets, int32_t* hit_idx, float* hit_integral,
                      float* hit_dist, int64_t* stats)
{
    const long n_nodes = n_leaves - 1;
    const long n_packets = (n_rays + 31) / 32;
    int max_depth = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : max_depth)
    for (long p = 0; p < n_packets; ++p) {
        const long r0 = p * 32;
        const int nl = (int)((n_rays - r0) < 32 ? (n_rays - r0) : 32);
        const orc_ray* rays = (const orc_ray*)rays7 + r0;
        float invd[32][3], org[32][3];
        int32_t cnt[32];
        float cum[32];
        int32_t cur[32];
        for (int l = 0; l < nl; ++l) {
            invd[l][0] = 1.0f / rays[l].dx;
            invd[l][1] = 1.0f / rays[l].dy;
            invd[l][2] = 1.0f / rays[l].dz;
            org[l][0] = rays[l].ox; org[l][1] = rays[l].oy; org[l][2] = rays[l].oz;
            cnt[l] = 0; cum[l] = 0.0f;
            cur[l] = (mode == 2) ? offsets[r0 + l] : 0;
        }
        int64_t nv = 0, lv = 0, ps = 0;
        int32_t stack[4096];
        int sp = 0;
        stack[sp++] = root;
        while (sp > 0) {
            if (sp > max_depth) max_depth = sp;
            int32_t idx = stack[--sp];
            if (idx < n_nodes) {
                const int32_t* node = nodes + 16 * (long)idx;
                int anyL = 0, anyR = 0;
                for (int l = 0; l < nl; ++l) {
                    int h = aabbs_hit(invd[l], org[l], rays[l].length, node);
                    anyR |= h & 1;
                    anyL |= h >> 1;
                }
                ++nv;
                if (anyR) stack[sp++] = node[1];
                if (anyL) stack[sp++] = node[0];
            } else {
                const int32_t* leaf = leaves + 4 * (long)(idx - n_nodes);
                const int first = leaf[0], count = leaf[1];
                ++lv; ps += count;
                for (int i = 0; i < count; ++i) {
                    const float* s = s4 + 4 * (long)(first + i);
                    for (int l = 0; l < nl; ++l) {
                        float b2, dot;
                        if (sphere_hit(&rays[l], s, &b2, &dot)) {
                            if (mode == 0) cnt[l]++;
                            else if (mode == 1) cum[l] = kernel_accumulate(cum[l], b2, s[3]);
                            else {
                                int32_t w = cur[l]++;
                                hit_idx[w] = first + i;
                                hit_integral[w] = kernel_integral(b2, s[3]);
                                hit_dist[w] = dot;
                            }
                        }
                    }
                }
            }
        }
        for (int l = 0; l < nl; ++l) {
            if (mode == 0) out_counts[r0 + l] = cnt[l];
            else if (mode == 1) out_cum[r0 + l] = cum[l];
        }
        if (stats) { stats[3 * p] = nv; stats[3 * p + 1] = lv; stats[3 * p + 2] = ps; }
    }
    return max_depth;
}

/* Host brute force: the pattern of tests/tree_traversal/tree_traversal.cu:65-79
 * (every ray against every sphere), with the device FMA form of sphere_hit.
 * mode 0 -> counts, mode 1 -> cumulative column density (sphere index order). */
ORC_API void orc_brute(const float* rays7, long n_rays, const float* s4, long n,
                       int mode, int32_t* out_counts, float* out_cum)
{
    const orc_ray* rays = (const orc_ray*)rays7;
#pragma omp parallel for schedule(dynamic, 4)
    for (long r = 0; r < n_rays; ++r) {
        int32_t c = 0;
        float cum = 0.0f;
        for (long i = 0; i < n; ++i) {
            float b2, dot;
            if (sphere_hit(&rays[r], s4 + 4 * i, &b2, &dot)) {
                if (mode == 0) ++c;
                else cum = kernel_accumulate(cum, b2, s4[4 * i + 3]);
            }
        }
        if (mode == 0) out_counts[r] = c; else out_cum[r] = cum;
    }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Per-ray stable sort of hits by distance: cuda/sort.cuh:100-131 (sgpu        */
/* SegSortPairsFromIndices is a stable segmented sort with less<float>, then   */
/* indices and one payload are gathered through the same map).                 */
/* ------------------------------------------------------------------------- */
static void merge_sort_idx(const float* key, int32_t* idx, int32_t* tmp, long n)
{
    for (long w = 1; w < n; w *= 2) {
        for (long lo = 0; lo < n; lo += 2 * w) {
            long mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            long a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (key[idx[b]] < key[idx[a]]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, sizeof(int32_t) * (size_t)n);
    }
}

ORC_API void orc_sort_by_distance(float* dist, const int32_t* offsets, long n_rays,
                                  long total, int32_t* hit_idx, float* hit_data)
{
#pragma omp parallel for schedule(dynamic, 8)
    for (long r = 0; r < n_rays; ++r) {
        long b = offsets[r], e = (r + 1 < n_rays) ? offsets[r + 1] : total;
        long n = e - b;
        if (n <= 1) continue;
        int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * 2);
        float* t = (float*)malloc(sizeof(float) * (size_t)n);
        for (long i = 0; i < n; ++i) idx[i] = (int32_t)i;
        merge_sort_idx(dist + b, idx, idx + n, n);
        for (long i = 0; i < n; ++i) t[i] = dist[b + idx[i]];
        memcpy(dist + b, t, sizeof(float) * (size_t)n);
        for (long i = 0; i < n; ++i) t[i] = hit_data[b + idx[i]];
        memcpy(hit_data + b, t, sizeof(float) * (size_t)n);
        int32_t* ti = (int32_t*)t;
        for (long i = 0; i < n; ++i) ti[i] = hit_idx[b + idx[i]];
        memcpy(hit_idx + b, ti, sizeof(int32_t) * (size_t)n);
        free(idx); free(t);
    }
}

/* ------------------------------------------------------------------------- */
/* Deterministic ray generators.                                              */
/* ------------------------------------------------------------------------- */

/* cuda/kernels/gen_rays.cuh:46-65: invR = (float) rnorm3d((double)dx,..) then
 * three float multiplies.  libdevice's rnorm3d is restated as 1/sqrt in double
 * (agrees to the last bit of the float result except in rare double-rounding
 * cases; tests allow 1 ulp). */
static inline float normalize_dir(float dx, float dy, float dz, orc_ray* r)
{
    double x = dx, y = dy, z = dz;
    float invR = (float)(1.0 / sqrt(x * x + y * y + z * z));
    r->dx = dx * invR; r->dy = dy * invR; r->dz = dz * invR;
    return invR;
}

/* one_to_many_rays_kernel, cuda/kernels/gen_rays.cuh:209-244 (NoSort order). */
ORC_API void orc_one_to_many_rays(const float* o3, const float* points, int stride,
                                  long n, float* rays7)
{
    orc_ray* rays = (orc_ray*)rays7;
    for (long i = 0; i < n; ++i) {
        const float* p = points + (long)stride * i;
        float dx = p[0] - o3[0], dy = p[1] - o3[1], dz = p[2] - o3[2];
        float invR = normalize_dir(dx, dy, dz, &rays[i]);
        rays[i].ox = o3[0]; rays[i].oy = o3[1]; rays[i].oz = o3[2];
        rays[i].length = (float)(1.0 / (double)invR);
    }
}

/* ray_dir_morton_key, cuda/kernels/gen_rays.cuh:38-43 with
 * morton_key(float,float,float), generic/morton.h:32-42. */
ORC_API uint32_t orc_ray_dir_key(const float* ray7)
{
    float x = (ray7[0] + 1.0f) / 2.0f, y = (ray7[1] + 1.0f) / 2.0f, z = (ray7[2] + 1.0f) / 2.0f;
    return orc_morton_key30(cvt_rzi_u32(1023.0f * x), cvt_rzi_u32(1023.0f * y),
                            cvt_rzi_u32(1023.0f * z));
}

/* HEALPix NESTED pixel -> unit vector:
 * RayVectorGeneration/src/chealpix/chealpix.c:112-126 (nest2xyf), :357-391
 * (pix2ang_nest_z_phi), :459-467 (pix2vec_nest). */
/* nest2xyf (chealpix.c:112-126): the ctab lookups de-interleave the pixel
 * index, ix bit k = pix bit 2k (ctab[m] moves even bits of m to bits 0..3 and
 * odd bits to bits 8..11, chealpix.c:76-79). */
static int compress_bits(int v)
{
    int out = 0;
    for (int b = 0; b < 16; ++b) out |= ((v >> (2 * b)) & 1) << b;
    return out;
}

ORC_API void orc_pix2vec_nest(long nside, long ipix, double* vec)
{
    static const int jrll[] = { 2,2,2,2,3,3,3,3,4,4,4,4 };
    static const int jpll[] = { 1,3,5,7,0,2,4,6,1,3,5,7 };
    const double halfpi = 1.570796326794896619231321691639751442099;
    int nside_ = (int)nside, pix = (int)ipix;
    int npface = nside_ * nside_;
    int face_num = pix / npface;
    pix &= (npface - 1);
    int ix = compress_bits(pix);
    int iy = compress_bits(pix >> 1);
    int nl4 = nside_ * 4;
    int npix_ = 12 * nside_ * nside_;
    double fact2_ = 4. / npix_;
    int nr, kshift;
    double z;
    int jr = (jrll[face_num] * nside_) - ix - iy - 1;
    if (jr < nside_) { nr = jr; z = 1 - nr * nr * fact2_; kshift = 0; }
    else if (jr > 3 * nside_) { nr = nl4 - jr; z = nr * nr * fact2_ - 1; kshift = 0; }
    else {
        double fact1_ = (nside_ << 1) * fact2_;
        nr = nside_; z = (2 * nside_ - jr) * fact1_; kshift = (jr - nside_) & 1;
    }
    int jp = (jpll[face_num] * nr + ix - iy + 1 + kshift) / 2;
    if (jp > nl4) jp -= nl4;
    if (jp < 1) jp += nl4;
    double phi = (jp - (kshift + 1) * 0.5) * (halfpi / nr);
    double stheta = sqrt((1. - z) * (1. + z));
    vec[0] = stheta * cos(phi);
    vec[1] = stheta * sin(phi);
    vec[2] = z;
}
