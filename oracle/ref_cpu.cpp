// ref_cpu.cpp -- the reference's own host-callable code behind a C interface.
//
// TEST INFRASTRUCTURE: compiled by oracle/build_ref.sh into oracle/_ref/libgrace_ref_cpu.so
// against the headers under /root/reference/include (nothing is copied).  It calls
// grace::sphere_hit (generic/intersect.h:10-55), grace::lerp (generic/interpolate.h:11-39,
// host branch) and grace::morton_key (generic/morton.h:14-29) exactly as the reference's
// host-side tests do (tests/tree_traversal/tree_traversal.cu:65-79,
// tests/morton_key_kernel/30bit_keys.cu:48-52).  Used to pin the oracle and as the
// "reference CPU" arm of bench.py.
#include "grace/generic/intersect.h"
#include "grace/generic/interpolate.h"
#include "grace/generic/morton.h"
#include "grace/ray.h"

#include <omp.h>
#include <stdint.h>

#define API extern "C" __attribute__((visibility("default")))

// numeric data of cuda/trace_sph.cuh:32-48 (that header needs Thrust + nvcc)
static const double table[51] = {
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

API int ref_num_threads(void) { return omp_get_max_threads(); }

API int ref_sphere_hit(const grace::Ray* ray, const float4* s, float* b2, float* dot)
{
    return grace::sphere_hit(*ray, *s, *b2, *dot) ? 1 : 0;
}

API uint32_t ref_morton_key30(uint32_t x, uint32_t y, uint32_t z) { return grace::morton_key(x, y, z); }
API uint64_t ref_morton_key63(uint64_t x, uint64_t y, uint64_t z) { return grace::morton_key(x, y, z); }

// tests/tree_traversal/tree_traversal.cu:65-79
API void ref_brute_hitcounts(const grace::Ray* rays, long n_rays, const float4* spheres, long n, int* out)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (long ri = 0; ri < n_rays; ++ri) {
        grace::Ray ray = rays[ri];
        int hits = 0;
        float b2, d;
        for (long si = 0; si < n; ++si)
            if (grace::sphere_hit(ray, spheres[si], b2, d)) ++hits;
        out[ri] = hits;
    }
}

// same loop with OnHit_sphere_cumulate's arithmetic (cuda/functors/trace.cuh:183-191)
API void ref_brute_cumulative(const grace::Ray* rays, long n_rays, const float4* spheres, long n, float* out)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (long ri = 0; ri < n_rays; ++ri) {
        grace::Ray ray = rays[ri];
        float cum = 0.f, b2, d;
        for (long si = 0; si < n; ++si) {
            const float4 s = spheres[si];
            if (grace::sphere_hit(ray, s, b2, d)) {
                float ir = 1.f / s.w;
                float b = (51 - 1) * (sqrtf(b2) * ir);
                float integral = grace::lerp(b, table, 51);
                integral *= (ir * ir);
                cum += integral;
            }
        }
        out[ri] = cum;
    }
}
