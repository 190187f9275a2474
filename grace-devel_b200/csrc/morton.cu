// morton.cu -- fused centroid bounds reduction and 30/63-bit Morton keys.
//
// Reference behaviour (GRACE): cuda/kernels/aabb.cuh:14-49 (centroids),
// cuda/util/extrema.cuh:502-513,667-678 (min_vec3/max_vec3 via two thrust::reduce
// passes over a temporary float3 array + host read-backs), cuda/kernels/morton.cuh:30-55
// (keys kernel, grid capped at 112 blocks), :107-113 (scale = span/(top-bot) on the host).
//
// B200 design: one pass over the float4 spheres computes min and max of all four
// components with no temporary centroid array (HBM-bound: 16 B/particle read);
// per-block partials are folded by the last block to finish (ticket), so there is
// no host round trip.  The key kernel reads the bounds from device memory and
// derives the scale itself with the same IEEE float division the host would do.
#include "common.cuh"
#include "morton_device.cuh"

#include <math_constants.h>

namespace {

constexpr int MM_THREADS = 256;

__device__ __forceinline__ void warp_minmax(float (&lo)[4], float (&hi)[4])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
}

// partials: [gridDim.x][8] floats {min xyzw, max xyzw}
__global__ void __launch_bounds__(MM_THREADS)
minmax_kernel(const float4* __restrict__ s, size_t n, float* __restrict__ partials,
              int* __restrict__ ticket, float* __restrict__ out, int n_out_comp)
{
    float lo[4] = { CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F };
    float hi[4] = { -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F };
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float4 v = gb_ld_stream_f4(s + i);
        lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
        lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
        lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
        lo[3] = fminf(lo[3], v.w); hi[3] = fmaxf(hi[3], v.w);
    }
    __shared__ float sm[MM_THREADS / 32][8];
    __shared__ bool is_last;
    warp_minmax(lo, hi);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { sm[w][k] = lo[k]; sm[w][4 + k] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = sm[0][threadIdx.x];
        for (int i = 1; i < MM_THREADS / 32; ++i)
            v = threadIdx.x < 4 ? fminf(v, sm[i][threadIdx.x]) : fmaxf(v, sm[i][threadIdx.x]);
        partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = atomicAdd(ticket, 1);
        is_last = (t == (int)gridDim.x - 1);
        if (is_last) *ticket = 0;   // self-reset for the next call
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // Last block: fold all partials.  Thread t handles component t & 7.
    const int comp = threadIdx.x & 7;
    const bool is_min = comp < 4;
    float v = is_min ? CUDART_INF_F : -CUDART_INF_F;
    for (int b = threadIdx.x >> 3; b < (int)gridDim.x; b += MM_THREADS / 8) {
        float p = __ldcg(partials + (size_t)b * 8 + comp);
        v = is_min ? fminf(v, p) : fmaxf(v, p);
    }
    // reduce across the 32 threads sharing a component (stride 8 in the block)
    __shared__ float red[MM_THREADS];
    red[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < 8) {
        for (int i = threadIdx.x + 8; i < MM_THREADS; i += 8)
            v = is_min ? fminf(v, red[i]) : fmaxf(v, red[i]);
        // n_out_comp = 3 -> {min xyz, max xyz}; 4 -> {min xyzw, max xyzw}
        const int k = comp & 3;
        if (k < n_out_comp) out[(is_min ? 0 : n_out_comp) + k] = v;
    }
}

struct Bounds6 { float v[6]; };

// Bounds come either from device memory (d_bounds6 != NULL; no host round trip after
// the bounds kernel) or by value (explicit host bounds).
// FUSED (the keys + sort entry point): the kernel also (a) counts the digits of every radix pass
// -- the onesweep sort needs them up front and the keys are in registers here -- and (b) writes
// a copy of the spheres for the final gather to read, so that the gather can write the sorted
// spheres straight into the caller's array.
template <typename KeyT, bool FUSED>
__global__ void __launch_bounds__(256)
morton_keys_kernel(const float4* __restrict__ s, size_t n, const float* __restrict__ d_bounds6,
                   const Bounds6 hb, KeyT* __restrict__ keys, uint32_t* __restrict__ hist, int passes,
                   float4* __restrict__ copy)
{
    __shared__ uint32_t sh[FUSED ? 8 * 256 : 1];
    if (FUSED) {
        for (int i = threadIdx.x; i < passes * 256; i += 256) sh[i] = 0;
        __syncthreads();
    }
    const float* bounds6 = d_bounds6 ? d_bounds6 : hb.v;
    const float3 bot = make_float3(bounds6[0], bounds6[1], bounds6[2]);
    const float3 scale = gb_morton_scale<KeyT>(bot, make_float3(bounds6[3], bounds6[4], bounds6[5]));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // whole warps iterate together (match_all below needs every lane)
    const size_t n_up = (n + 31) / 32 * 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
        const bool live = i < n;
        KeyT k = 0;
        if (live) {
            const float4 v = gb_ld_stream_f4(s + i);
            k = gb_morton_key<KeyT>(v, bot, scale);
            keys[i] = k;
            if (FUSED) copy[i] = v;
        }
        if (FUSED) {
            for (int p = 0; p < passes; ++p) {
                const unsigned d = (unsigned)(k >> (8 * p)) & 255u;
                int all_same;
                __match_all_sync(0xffffffffu, live ? d : 0x100u + (threadIdx.x & 31), &all_same);
                if (all_same) { if ((threadIdx.x & 31) == 0) atomicAdd(&sh[p * 256 + d], 32u); }
                else if (live) atomicAdd(&sh[p * 256 + d], 1u);
            }
        }
    }
    if (FUSED) {
        __syncthreads();
        for (int i = threadIdx.x; i < passes * 256; i += 256)
            if (sh[i]) atomicAdd(&hist[i], sh[i]);
    }
}

int grid_for(const grace_b200_ctx* ctx, size_t n, int threads, int per_sm)
{
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int run_minmax(grace_b200_ctx* ctx, const float* d_s, size_t n, float* d_out, int ncomp, void* stream)
{
    GB_REQUIRE(ctx && d_s && d_out, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n > 0, GRACE_B200_EINVAL, "bounds of an empty particle set");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(ctx, n, MM_THREADS * 4, 8);
    float* partials = (float*)gb_workspace(ctx, (size_t)grid * 8 * sizeof(float));
    if (!partials) return GRACE_B200_ENOMEM;
    minmax_kernel<<<grid, MM_THREADS, 0, st>>>((const float4*)d_s, n, partials,
                                               ctx->d_scalars + GB_SC_TICKET0, d_out, ncomp);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // namespace

template <typename KeyT>
int gb_launch_morton_keys(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                          const float* d_bounds6, const float* h_bounds6, KeyT* d_keys,
                          cudaStream_t st)
{
    if (n == 0) return GRACE_B200_OK;
    Bounds6 hb = {};
    if (!d_bounds6) for (int i = 0; i < 6; ++i) hb.v[i] = h_bounds6[i];
    morton_keys_kernel<KeyT, false><<<grid_for(ctx, n, 256 * 2, 16), 256, 0, st>>>(
        (const float4*)d_spheres4, n, d_bounds6, hb, d_keys, nullptr, 0, nullptr);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

// keys + digit histograms of all radix passes + a copy of the spheres (see the kernel)
template <typename KeyT>
int gb_launch_morton_keys_fused(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, const float* h_bounds6, KeyT* d_keys,
                                uint32_t* d_hist, int passes, float* d_copy4, cudaStream_t st)
{
    GB_REQUIRE(d_bounds6 || h_bounds6, GRACE_B200_EINVAL, "bounds missing");
    if (n == 0) return GRACE_B200_OK;
    Bounds6 hb = {};
    if (!d_bounds6) for (int i = 0; i < 6; ++i) hb.v[i] = h_bounds6[i];
    morton_keys_kernel<KeyT, true><<<grid_for(ctx, n, 256 * 2, 8), 256, 0, st>>>(
        (const float4*)d_spheres4, n, d_bounds6, hb, d_keys, d_hist, passes, (float4*)d_copy4);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}
template int gb_launch_morton_keys_fused<uint32_t>(grace_b200_ctx*, const float*, size_t, const float*,
                                                   const float*, uint32_t*, uint32_t*, int, float*, cudaStream_t);
template int gb_launch_morton_keys_fused<uint64_t>(grace_b200_ctx*, const float*, size_t, const float*,
                                                   const float*, uint64_t*, uint32_t*, int, float*, cudaStream_t);
template int gb_launch_morton_keys<uint32_t>(grace_b200_ctx*, const float*, size_t, const float*,
                                             const float*, uint32_t*, cudaStream_t);
template int gb_launch_morton_keys<uint64_t>(grace_b200_ctx*, const float*, size_t, const float*,
                                             const float*, uint64_t*, cudaStream_t);

extern "C" {

int grace_b200_bounds_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                         float* d_bounds6, void* stream)
{
    return run_minmax(ctx, d_spheres4, n, d_bounds6, 3, stream);
}

int grace_b200_minmax_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                         float* d_minmax8, void* stream)
{
    return run_minmax(ctx, d_spheres4, n, d_minmax8, 4, stream);
}

int grace_b200_minmax_f4_host(grace_b200_ctx* ctx, const float* d_spheres4, size_t n, float* h_minmax8, void* stream)
{
    GB_REQUIRE(ctx && h_minmax8, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    // result through the context's device scalars and pinned mirror: no allocation per call
    float* d_out = (float*)(ctx->d_scalars + GB_SC_MINMAX);
    int rc = run_minmax(ctx, d_spheres4, n, d_out, 4, stream);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_MINMAX, d_out, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) h_minmax8[i] = ((const float*)(ctx->h_pinned + GB_SC_MINMAX))[i];
    return GRACE_B200_OK;
}

int grace_b200_morton_keys30_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, uint32_t* d_keys, void* stream)
{
    GB_REQUIRE(ctx && d_spheres4 && d_bounds6 && d_keys, GRACE_B200_EINVAL, "NULL argument");
    return gb_launch_morton_keys<uint32_t>(ctx, d_spheres4, n, d_bounds6, nullptr, d_keys,
                                           (cudaStream_t)stream);
}

int grace_b200_morton_keys63_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, uint64_t* d_keys, void* stream)
{
    GB_REQUIRE(ctx && d_spheres4 && d_bounds6 && d_keys, GRACE_B200_EINVAL, "NULL argument");
    return gb_launch_morton_keys<uint64_t>(ctx, d_spheres4, n, d_bounds6, nullptr, d_keys,
                                           (cudaStream_t)stream);
}

} // extern "C"
