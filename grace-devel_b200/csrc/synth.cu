// synth.cu -- synthetic "Gadget-shaped" SPH snapshot generated on the device.
//
// The reference ships no data file (its profilers default to a 128^3 Gadget-2 snapshot
// that is not in the repository, tests/profile_tree_gadget/profile_tree_gadget.cu:35), so
// the benchmark inputs are synthetic: float4 {x,y,z,h} in [0,1)^3, 30 %
// uniform background + 70 % in 512 Plummer-profile halos (centres uniform, masses from a
// power law, scale radius set from the mass so that each halo's central density is
// 10^3..10^4 times the mean -- the density contrast of a cosmological gas snapshot --
// and clamped to [0.002, 0.03]), smoothing length from the
// analytic local number density with N_ngb = 32, clamped to [1e-5, 0.1], particles stored
// in Peano-Hilbert cell order at 2^7 cells per side (Gadget's on-disk order).
// Counter-based hashing (Wang/Jenkins, the hash of tests/helper/random.cuh:20-29) makes
// particle i a pure function of (seed, i).
#include "common.cuh"
#include "radix_sort.cuh"

namespace {

constexpr int N_HALOS = 512;
constexpr float F_BACKGROUND = 0.3f;

__host__ __device__ inline unsigned wang_hash(unsigned a)
{
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}

__host__ __device__ inline float u01(unsigned seed, unsigned long long i, unsigned k)
{
    unsigned h = wang_hash((unsigned)i * 9u + k + 0x9e3779b9u * (seed + 1u));
    h = wang_hash(h ^ (unsigned)(i >> 32) ^ (k * 0x85ebca6bu));
    return (float)(h >> 8) * (1.0f / 16777216.0f);      // [0, 1)
}

struct Halo { float cx, cy, cz, a, cum_mass, mass; };

__global__ void make_halos_kernel(Halo* halos, unsigned seed)
{
    // single block of N_HALOS threads
    __shared__ float m[N_HALOS];
    const int h = threadIdx.x;
    Halo H;
    H.cx = u01(seed ^ 0xabcdu, h, 0);
    H.cy = u01(seed ^ 0xabcdu, h, 1);
    H.cz = u01(seed ^ 0xabcdu, h, 2);
    const float um = u01(seed ^ 0xabcdu, h, 4);
    const float mass = powf(1.0f - um * 0.999f, -1.0f / 0.9f);        // power law dN/dm ~ m^-1.9
    m[h] = mass;
    __syncthreads();
    float tot = 0.f, cum = 0.f;
    for (int i = 0; i < N_HALOS; ++i) { tot += m[i]; if (i <= h) cum += m[i]; }
    H.mass = mass / tot;
    H.cum_mass = cum / tot;
    // central Plummer density 3 M / (4 pi a^3) = contrast * mean density, contrast log-uniform
    // in [1e3, 1e4]; M = (1 - F_BACKGROUND) * mass fraction (in units of the total mass)
    const float contrast = 1.0e3f * powf(10.0f, u01(seed ^ 0xabcdu, h, 3));
    H.a = cbrtf(3.0f * (1.0f - F_BACKGROUND) * H.mass / (4.0f * 3.14159265f * contrast));
    H.a = fminf(fmaxf(H.a, 0.002f), 0.03f);
    halos[h] = H;
}

__device__ __forceinline__ float wrap01(float x)
{
    x = x - floorf(x);
    return fminf(x, 0.99999994f);
}

__global__ void __launch_bounds__(256)
positions_kernel(float4* __restrict__ out, size_t n, const Halo* __restrict__ halos_g, unsigned seed)
{
    __shared__ Halo halos[N_HALOS];
    for (int i = threadIdx.x; i < N_HALOS; i += blockDim.x) halos[i] = halos_g[i];
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x, y, z;
        if (u01(seed, i, 0) < F_BACKGROUND) {
            x = u01(seed, i, 1); y = u01(seed, i, 2); z = u01(seed, i, 3);
        } else {
            const float um = u01(seed, i, 1);
            int lo = 0, hi = N_HALOS - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (halos[mid].cum_mass < um) lo = mid + 1; else hi = mid; }
            const Halo H = halos[lo];
            // Plummer: M(<r)/M = r^3/(r^2+a^2)^{3/2}; invert, truncated at ~5a
            const float u = fmaxf(u01(seed, i, 2) * 0.94f, 1e-6f);
            const float r = H.a * rsqrtf(powf(u, -2.0f / 3.0f) - 1.0f);
            const float ct = 2.0f * u01(seed, i, 3) - 1.0f;
            const float sn = sqrtf(fmaxf(0.f, 1.0f - ct * ct));
            const float ph = 6.2831853f * u01(seed, i, 4);
            x = wrap01(H.cx + r * sn * cosf(ph));
            y = wrap01(H.cy + r * sn * sinf(ph));
            z = wrap01(H.cz + r * ct);
        }
        out[i] = make_float4(x, y, z, 0.f);
    }
}

__global__ void __launch_bounds__(256)
smoothing_kernel(float4* __restrict__ s, size_t n, const Halo* __restrict__ halos_g)
{
    __shared__ Halo halos[N_HALOS];
    for (int i = threadIdx.x; i < N_HALOS; i += blockDim.x) halos[i] = halos_g[i];
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float fn = (float)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 p = s[i];
        float dens = F_BACKGROUND * fn;     // number density of the uniform component
        for (int h = 0; h < N_HALOS; ++h) {
            const Halo H = halos[h];
            float dx = fabsf(p.x - H.cx), dy = fabsf(p.y - H.cy), dz = fabsf(p.z - H.cz);
            dx = fminf(dx, 1.f - dx); dy = fminf(dy, 1.f - dy); dz = fminf(dz, 1.f - dz);
            const float q = 1.0f + (dx * dx + dy * dy + dz * dz) / (H.a * H.a);
            // Plummer density 3M/(4 pi a^3) (1 + r^2/a^2)^{-5/2}
            dens += (1.0f - F_BACKGROUND) * fn * H.mass * 0.238732415f / (H.a * H.a * H.a) *
                    rsqrtf(q * q * q * q * q);
        }
        float h = cbrtf(3.0f * 32.0f / (4.0f * 3.14159265f * dens));
        p.w = fminf(fmaxf(h, 1e-5f), 0.1f);
        s[i] = p;
    }
}

// 3-D Hilbert (Peano-Hilbert) key, `bits` bits per axis (Skilling's transpose method).
__device__ __forceinline__ unsigned hilbert_key(unsigned x, unsigned y, unsigned z, int bits)
{
    unsigned X[3] = { x, y, z };
    const unsigned M = 1u << (bits - 1);
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    unsigned key = 0;
    for (int b = bits - 1; b >= 0; --b)
#pragma unroll
        for (int i = 0; i < 3; ++i) key = (key << 1) | ((X[i] >> b) & 1u);
    return key;
}

__global__ void __launch_bounds__(256)
ph_keys_kernel(const float4* __restrict__ s, size_t n, uint32_t* __restrict__ keys)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 p = s[i];
        const unsigned cx = min(127u, (unsigned)(p.x * 128.0f));
        const unsigned cy = min(127u, (unsigned)(p.y * 128.0f));
        const unsigned cz = min(127u, (unsigned)(p.z * 128.0f));
        keys[i] = hilbert_key(cx, cy, cz, 7);
    }
}

} // namespace

extern "C" int grace_b200_synth_gadget_f4(grace_b200_ctx* ctx, float* d_spheres4, size_t n,
                                          unsigned int seed, void* stream)
{
    GB_REQUIRE(ctx && (d_spheres4 || n == 0), GRACE_B200_EINVAL, "NULL argument");
    if (n == 0) return GRACE_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sort_ws = gb_sort_workspace_bytes(n, 4);
    const size_t bytes = sort_ws + 2 * gb_align(n * 4) + gb_align(n * 16) + gb_align(sizeof(Halo) * N_HALOS) + 256;
    char* ws = (char*)gb_workspace(ctx, bytes);
    if (!ws) return GRACE_B200_ENOMEM;
    char* p = ws + sort_ws;
    uint32_t* keys = (uint32_t*)p; p += gb_align(n * 4);
    uint32_t* perm = (uint32_t*)p; p += gb_align(n * 4);
    float4* tmp = (float4*)p; p += gb_align(n * 16);
    Halo* halos = (Halo*)p;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    make_halos_kernel<<<1, N_HALOS, 0, st>>>(halos, seed);
    GB_LAUNCH_CHECK();
    positions_kernel<<<(int)blocks, 256, 0, st>>>(tmp, n, halos, seed);
    GB_LAUNCH_CHECK();
    smoothing_kernel<<<(int)blocks, 256, 0, st>>>(tmp, n, halos);
    GB_LAUNCH_CHECK();
    ph_keys_kernel<<<(int)blocks, 256, 0, st>>>(tmp, n, keys);
    GB_LAUNCH_CHECK();
    int rc = gb_sort_pairs<uint32_t>(ctx, keys, keys, perm, n, 21, ws, nullptr, st);
    if (rc) return rc;
    return gb_gather_records(tmp, d_spheres4, perm, n, 16, ctx->sm_count, st);
}
