// rays.cu -- ray generators.
//
// Reference behaviour (GRACE): kernels cuda/kernels/gen_rays.cuh:104-395, wrappers
// :416-787, public API cuda/gen_rays.cuh:26-399.  Random generators use the cuRAND
// device API (XORWOW, one sub-sequence per thread); the number of generator states is
// min(3*SMs*128, N_rays)+511 (the rounding expression at :434-436 evaluates to that),
// so results depend on the SM count by construction (cuda/gen_rays.cuh:21-24).  Sorted
// variants sort 28-byte Ray structs with thrust::sort_by_key.
//
// B200 design: the PRNG state never round-trips through global memory (curand_init and
// generation are fused in one kernel, same sub-sequences => same numbers); rays are
// written once, unsorted, to the workspace and placed by ONE gather after the onesweep
// (key, index) sort.  Floating-point expressions follow the contraction nvcc applies
// to the reference (SASS-verified): c = fma(x, v, y*u) + n; w = fma(fma(i+1, dw, -(i*dw)),
// rand, i*dw); direction keys (d+1)*0.5*1023 truncated.
#include "common.cuh"
#include "morton_device.cuh"
#include "radix_sort.cuh"

#include <curand_kernel.h>

#include <cmath>

namespace {

constexpr int RAYS_THREADS = 512;   // RAYS_THREADS_PER_BLOCK, cuda/kernel_config.h:10

struct F3 { float x, y, z; };

// cuda/kernels/gen_rays.cuh:46-65: double-precision rnorm3d, then three float multiplies.
__device__ __forceinline__ float normalise_dir(float dx, float dy, float dz, grace_b200_ray& ray)
{
    const float invR = (float)rnorm3d((double)dx, (double)dy, (double)dz);
    ray.dx = __fmul_rn(dx, invR);
    ray.dy = __fmul_rn(dy, invR);
    ray.dz = __fmul_rn(dz, invR);
    return invR;
}

// ray_dir_morton_key, cuda/kernels/gen_rays.cuh:38-43 + generic/morton.h:32-42.
__device__ __forceinline__ uint32_t dir_key(const grace_b200_ray& r)
{
    const uint32_t x = __float2uint_rz(__fmul_rn(__fmul_rn(__fadd_rn(r.dx, 1.0f), 0.5f), 1023.0f));
    const uint32_t y = __float2uint_rz(__fmul_rn(__fmul_rn(__fadd_rn(r.dy, 1.0f), 0.5f), 1023.0f));
    const uint32_t z = __float2uint_rz(__fmul_rn(__fmul_rn(__fadd_rn(r.dz, 1.0f), 0.5f), 1023.0f));
    return gb_interleave(x, y, z);
}

__device__ __forceinline__ void store_ray(grace_b200_ray* rays, size_t i, const grace_b200_ray& r)
{
    rays[i] = r;
}

// gen_uniform_rays_kernel / _single_octant_kernel, cuda/kernels/gen_rays.cuh:126-205,
// with init_PRNG_kernel (:104-118) fused: thread t uses sub-sequence t of `seed`.
__global__ void __launch_bounds__(RAYS_THREADS)
uniform_rays_kernel(grace_b200_ray* __restrict__ rays, uint32_t* __restrict__ keys, size_t n_rays,
                    float ox, float oy, float oz, float length, int octant,
                    unsigned long long seed)
{
    size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (tid >= n_rays) return;
    curandStateXORWOW_t state;
    curand_init(seed, tid, 0, &state);
    const float sx = (octant & 0x4) ? 1.f : -1.f;
    const float sy = (octant & 0x2) ? 1.f : -1.f;
    const float sz = (octant & 0x1) ? 1.f : -1.f;
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    for (; tid < n_rays; tid += stride) {
        float dx = curand_normal(&state);
        float dy = curand_normal(&state);
        float dz = curand_normal(&state);
        if (octant >= 0) {
            dx = __fmul_rn(sx, fabsf(dx));
            dy = __fmul_rn(sy, fabsf(dy));
            dz = __fmul_rn(sz, fabsf(dz));
        }
        grace_b200_ray ray;
        normalise_dir(dx, dy, dz, ray);
        keys[tid] = dir_key(ray);
        ray.ox = ox; ray.oy = oy; ray.oz = oz; ray.length = length;
        store_ray(rays, tid, ray);
    }
}

// one_to_many_rays_kernel, cuda/kernels/gen_rays.cuh:209-244.
__global__ void __launch_bounds__(256)
one_to_many_kernel(grace_b200_ray* __restrict__ rays, uint32_t* __restrict__ keys, size_t n_rays,
                   float ox, float oy, float oz, const float* __restrict__ points, int stride_f,
                   int key_mode, F3 bot, F3 top)
{
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    F3 scale = { 0, 0, 0 };
    if (key_mode == GRACE_B200_ENDPOINT_SORT) {
        const float3 s = gb_morton_scale<uint32_t>(make_float3(bot.x, bot.y, bot.z),
                                                   make_float3(top.x, top.y, top.z));
        scale = { s.x, s.y, s.z };
    }
    for (size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x; t < n_rays; t += stride) {
        const float* p = points + t * (size_t)stride_f;
        const float px = p[0], py = p[1], pz = p[2];
        const float dx = __fsub_rn(px, ox), dy = __fsub_rn(py, oy), dz = __fsub_rn(pz, oz);
        grace_b200_ray ray;
        const float invR = normalise_dir(dx, dy, dz, ray);
        ray.ox = ox; ray.oy = oy; ray.oz = oz;
        ray.length = (float)(1.0 / (double)invR);          // :229
        if (key_mode == GRACE_B200_DIRECTION_SORT) keys[t] = dir_key(ray);
        else if (key_mode == GRACE_B200_ENDPOINT_SORT)
            keys[t] = gb_morton_key<uint32_t>(make_float4(px, py, pz, 0.f),
                                              make_float3(bot.x, bot.y, bot.z),
                                              make_float3(scale.x, scale.y, scale.z));
        store_ray(rays, t, ray);
    }
}

// plane_parallel_random_rays_kernel, cuda/kernels/gen_rays.cuh:247-316.
__global__ void __launch_bounds__(RAYS_THREADS)
plane_parallel_kernel(grace_b200_ray* __restrict__ rays, int width, size_t n_rays, F3 base,
                      F3 dw, F3 dh, float length, F3 normal, unsigned long long seed)
{
    size_t tid = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (tid >= n_rays) return;
    curandStateXORWOW_t state;
    curand_init(seed, tid, 0, &state);
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    for (; tid < n_rays; tid += stride) {
        const int i = (int)(tid % (size_t)width), j = (int)(tid / (size_t)width);
        const float fi = (float)i, fi1 = (float)(i + 1), fj = (float)j, fj1 = (float)(j + 1);
        const float rand_w = curand_uniform(&state);
        const float rand_h = curand_uniform(&state);
        // zero_one_to_a_b (:68-76) as contracted: fma(fma(i+1, d, -(i*d)), rand, i*d)
        const float awx = __fmul_rn(fi, dw.x), awy = __fmul_rn(fi, dw.y), awz = __fmul_rn(fi, dw.z);
        const float ahx = __fmul_rn(fj, dh.x), ahy = __fmul_rn(fj, dh.y), ahz = __fmul_rn(fj, dh.z);
        const float wx = __fmaf_rn(__fmaf_rn(fi1, dw.x, -awx), rand_w, awx);
        const float wy = __fmaf_rn(__fmaf_rn(fi1, dw.y, -awy), rand_w, awy);
        const float wz = __fmaf_rn(__fmaf_rn(fi1, dw.z, -awz), rand_w, awz);
        const float hx = __fmaf_rn(__fmaf_rn(fj1, dh.x, -ahx), rand_h, ahx);
        const float hy = __fmaf_rn(__fmaf_rn(fj1, dh.y, -ahy), rand_h, ahy);
        const float hz = __fmaf_rn(__fmaf_rn(fj1, dh.z, -ahz), rand_h, ahz);
        grace_b200_ray ray;
        ray.dx = normal.x; ray.dy = normal.y; ray.dz = normal.z;
        ray.ox = __fadd_rn(__fadd_rn(base.x, wx), hx);
        ray.oy = __fadd_rn(__fadd_rn(base.y, wy), hy);
        ray.oz = __fadd_rn(__fadd_rn(base.z, wz), hz);
        ray.length = length;
        store_ray(rays, tid, ray);
    }
}

// image_plane_coord, cuda/kernels/gen_rays.cuh:78-97, as contracted by nvcc:
// x = fma((i+.5)/rx, 2, -1) * aspect; y = 1 - 2*((j+.5)/ry); c = fma(x, v, y*u) + n
__device__ __forceinline__ F3 image_coord(int i, int j, F3 v, F3 u, F3 n, int rx, int ry, float aspect)
{
    const float tx = __fdiv_rn(__fadd_rn((float)i, 0.5f), (float)rx);
    const float ty = __fdiv_rn(__fadd_rn((float)j, 0.5f), (float)ry);
    const float x = __fmul_rn(__fmaf_rn(tx, 2.0f, -1.0f), aspect);
    const float y = __fsub_rn(1.0f, __fadd_rn(ty, ty));
    F3 c;
    c.x = __fadd_rn(__fmaf_rn(x, v.x, __fmul_rn(y, u.x)), n.x);
    c.y = __fadd_rn(__fmaf_rn(x, v.y, __fmul_rn(y, u.y)), n.y);
    c.z = __fadd_rn(__fmaf_rn(x, v.z, __fmul_rn(y, u.z)), n.z);
    return c;
}

// orthographic_projection_rays_kernel, cuda/kernels/gen_rays.cuh:319-360.
__global__ void __launch_bounds__(256)
ortho_kernel(grace_b200_ray* __restrict__ rays, int rx, int ry, size_t n_rays, F3 cam, F3 dir,
             F3 v, F3 u, float length)
{
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    const F3 zero = { 0.f, 0.f, 0.f };
    for (size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x; t < n_rays; t += stride) {
        const int i = (int)(t % (size_t)rx), j = (int)(t / (size_t)rx);
        const F3 c = image_coord(i, j, v, u, zero, rx, ry, 1.0f);
        grace_b200_ray ray;
        ray.dx = dir.x; ray.dy = dir.y; ray.dz = dir.z;
        ray.ox = __fadd_rn(cam.x, c.x); ray.oy = __fadd_rn(cam.y, c.y); ray.oz = __fadd_rn(cam.z, c.z);
        ray.length = length;
        store_ray(rays, t, ray);
    }
}

// perspective_projection_rays_kernel, cuda/kernels/gen_rays.cuh:362-395.
__global__ void __launch_bounds__(256)
pinhole_kernel(grace_b200_ray* __restrict__ rays, int rx, int ry, size_t n_rays, float aspect,
               F3 cam, F3 v, F3 u, F3 n, float length)
{
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    for (size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x; t < n_rays; t += stride) {
        const int i = (int)(t % (size_t)rx), j = (int)(t / (size_t)rx);
        const F3 c = image_coord(i, j, v, u, n, rx, ry, aspect);
        grace_b200_ray ray;
        normalise_dir(c.x, c.y, c.z, ray);
        ray.ox = cam.x; ray.oy = cam.y; ray.oz = cam.z;
        ray.length = length;
        store_ray(rays, t, ray);
    }
}

// HEALPix NESTED pixel -> unit vector, restating chealpix's nest2xyf (chealpix.c:112-126),
// pix2ang_nest_z_phi (:357-391) and pix2vec_nest (:459-467) in double precision.
__device__ __forceinline__ int compress_even_bits(unsigned v)
{
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0f0f0f0fu;
    v = (v | (v >> 4)) & 0x00ff00ffu;
    v = (v | (v >> 8)) & 0x0000ffffu;
    return (int)v;
}

__global__ void __launch_bounds__(256)
healpix_kernel(grace_b200_ray* __restrict__ rays, size_t n_rays, long nside, long first_pixel,
               float ox, float oy, float oz, float length)
{
    const int jrll[12] = { 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4 };
    const int jpll[12] = { 1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7 };
    const double halfpi = 1.570796326794896619231321691639751442099;
    const size_t stride = (size_t)blockDim.x * gridDim.x;
    const int ns = (int)nside;
    const int npface = ns * ns;
    const int nl4 = 4 * ns;
    const double fact2 = 4.0 / (12.0 * (double)ns * (double)ns);
    for (size_t t = threadIdx.x + (size_t)blockIdx.x * blockDim.x; t < n_rays; t += stride) {
        int pix = (int)(first_pixel + (long)t);
        const int face = pix / npface;
        pix &= (npface - 1);
        const int ix = compress_even_bits((unsigned)pix);
        const int iy = compress_even_bits((unsigned)pix >> 1);
        const int jr = jrll[face] * ns - ix - iy - 1;
        int nr, kshift;
        double z;
        if (jr < ns) { nr = jr; z = 1 - nr * nr * fact2; kshift = 0; }
        else if (jr > 3 * ns) { nr = nl4 - jr; z = nr * nr * fact2 - 1; kshift = 0; }
        else { const double fact1 = (ns << 1) * fact2; nr = ns; z = (2 * ns - jr) * fact1; kshift = (jr - ns) & 1; }
        int jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
        if (jp > nl4) jp -= nl4;
        if (jp < 1) jp += nl4;
        const double phi = (jp - (kshift + 1) * 0.5) * (halfpi / nr);
        const double st = sqrt((1.0 - z) * (1.0 + z));
        grace_b200_ray ray;
        ray.dx = (float)(st * cos(phi));
        ray.dy = (float)(st * sin(phi));
        ray.dz = (float)z;
        ray.ox = ox; ray.oy = oy; ray.oz = oz; ray.length = length;
        store_ray(rays, t, ray);
    }
}

int cap_blocks(const grace_b200_ctx* ctx, size_t n, int threads, int per_sm)
{
    size_t b = (n + threads - 1) / threads;
    const size_t cap = (size_t)ctx->sm_count * per_sm;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// init_PRNG's state count (cuda/kernels/gen_rays.cuh:428-438): the expression
// `factor * (N + factor - 1) / factor` evaluates left to right, i.e. N + factor - 1.
int prng_blocks(const grace_b200_ctx* ctx, size_t n_rays)
{
    long long N = 3LL * ctx->sm_count * 128;
    if ((long long)n_rays < N) N = (long long)n_rays;
    N = (long long)RAYS_THREADS * (N + RAYS_THREADS - 1) / RAYS_THREADS;
    return (int)(N / RAYS_THREADS);
}

// Sort the unsorted rays in `tmp` by `keys` (30-bit) into d_rays.
int sort_rays_into(grace_b200_ctx* ctx, char* ws, size_t sort_ws, uint32_t* keys, uint32_t* perm,
                   const grace_b200_ray* tmp, grace_b200_ray* d_rays, size_t n, cudaStream_t st)
{
    int rc = gb_sort_pairs<uint32_t>(ctx, keys, keys, perm, n, 32, ws, nullptr, st);
    if (rc) return rc;
    (void)sort_ws;
    return gb_gather_records(tmp, d_rays, perm, n, 28, ctx->sm_count, st);
}

struct RayWs { char* ws; size_t sort_ws; uint32_t* keys; uint32_t* perm; grace_b200_ray* tmp; };

bool ray_workspace(grace_b200_ctx* ctx, size_t n, RayWs& w)
{
    w.sort_ws = gb_sort_workspace_bytes(n, 4);
    const size_t bytes = w.sort_ws + 2 * gb_align(n * 4) + gb_align(n * 28) + 256;
    w.ws = (char*)gb_workspace(ctx, bytes);
    if (!w.ws) return false;
    char* p = w.ws + w.sort_ws;
    w.keys = (uint32_t*)p; p += gb_align(n * 4);
    w.perm = (uint32_t*)p; p += gb_align(n * 4);
    w.tmp = (grace_b200_ray*)p;
    return true;
}

// Host-side basis helpers: generic/vecmath.h:10-49 (host branch: float sum of
// squares, sqrt in float, 1./sqrt in double, component * double -> float).
F3 h_normalize3(F3 v)
{
    const double N = 1. / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    F3 r;
    r.x = (float)(v.x * N); r.y = (float)(v.y * N); r.z = (float)(v.z * N);
    return r;
}
F3 h_cross(F3 u, F3 v)
{
    volatile float a, b;
    F3 r;
    a = u.y * v.z; b = u.z * v.y; r.x = a - b;
    a = u.z * v.x; b = u.x * v.z; r.y = a - b;
    a = u.x * v.y; b = u.y * v.x; r.z = a - b;
    return r;
}
F3 h3(const float* p) { F3 r = { p[0], p[1], p[2] }; return r; }

} // namespace

extern "C" {

int grace_b200_uniform_random_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays,
                                   float ox, float oy, float oz, float length, int octant,
                                   unsigned long long seed, void* stream)
{
    GB_REQUIRE(ctx && (d_rays || n_rays == 0), GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(octant >= -1 && octant <= 7, GRACE_B200_EINVAL, "octant must be -1 or 0..7");
    if (n_rays == 0) return GRACE_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RayWs w;
    if (!ray_workspace(ctx, n_rays, w)) return GRACE_B200_ENOMEM;
    uniform_rays_kernel<<<prng_blocks(ctx, n_rays), RAYS_THREADS, 0, st>>>(
        w.tmp, w.keys, n_rays, ox, oy, oz, length, octant, seed);
    GB_LAUNCH_CHECK();
    return sort_rays_into(ctx, w.ws, w.sort_ws, w.keys, w.perm, w.tmp, d_rays, n_rays, st);
}

int grace_b200_one_to_many_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays,
                                float ox, float oy, float oz, const float* d_points,
                                int point_stride_floats, int sort_type, const float* h_bot3,
                                const float* h_top3, void* stream)
{
    GB_REQUIRE(ctx && (n_rays == 0 || (d_rays && d_points)), GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(point_stride_floats >= 3, GRACE_B200_EINVAL, "points need at least x, y, z");
    // cuda/gen_rays.cuh:124-130
    GB_REQUIRE(sort_type == GRACE_B200_NO_SORT || sort_type == GRACE_B200_DIRECTION_SORT ||
               sort_type == GRACE_B200_ENDPOINT_SORT, GRACE_B200_EINVAL, "Ray sort type not recognized");
    if (n_rays == 0) return GRACE_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = cap_blocks(ctx, n_rays, 256, 16);
    F3 bot = { 0, 0, 0 }, top = { 1, 1, 1 };
    if (sort_type == GRACE_B200_NO_SORT) {
        one_to_many_kernel<<<blocks, 256, 0, st>>>(d_rays, nullptr, n_rays, ox, oy, oz, d_points,
                                                   point_stride_floats, sort_type, bot, top);
        GB_LAUNCH_CHECK();
        return GRACE_B200_OK;
    }
    if (sort_type == GRACE_B200_ENDPOINT_SORT) {
        GB_REQUIRE((h_bot3 == nullptr) == (h_top3 == nullptr), GRACE_B200_EINVAL, "give both bounds or neither");
        if (h_bot3) { bot = h3(h_bot3); top = h3(h_top3); }
        else {
            // min_vec3/max_vec3 over the end points (cuda/gen_rays.cuh:117-122; the reference
            // passes AABB_bot twice there -- a bug that is not reproduced).
            GB_REQUIRE(point_stride_floats == 4, GRACE_B200_EINVAL,
                       "automatic end-point bounds need float4 points; pass bounds for other strides");
            float* d_b = (float*)gb_workspace(ctx, 4096);
            if (!d_b) return GRACE_B200_ENOMEM;
            d_b += 512;
            int rc = grace_b200_bounds_f4(ctx, d_points, n_rays, d_b, stream);
            if (rc) return rc;
            float hb[6];
            GB_CUDA(cudaMemcpyAsync(hb, d_b, sizeof(hb), cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaStreamSynchronize(st));
            bot = h3(hb); top = h3(hb + 3);
        }
    }
    RayWs w;
    if (!ray_workspace(ctx, n_rays, w)) return GRACE_B200_ENOMEM;
    one_to_many_kernel<<<blocks, 256, 0, st>>>(w.tmp, w.keys, n_rays, ox, oy, oz, d_points,
                                               point_stride_floats, sort_type, bot, top);
    GB_LAUNCH_CHECK();
    return sort_rays_into(ctx, w.ws, w.sort_ws, w.keys, w.perm, w.tmp, d_rays, n_rays, st);
}

int grace_b200_plane_parallel_random_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, int width,
                                          int height, const float* h_base3, const float* h_w3,
                                          const float* h_h3, float length, unsigned long long seed,
                                          void* stream)
{
    GB_REQUIRE(ctx && d_rays && h_base3 && h_w3 && h_h3, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(width > 0 && height > 0, GRACE_B200_EINVAL, "empty ray grid");
    const size_t n_rays = (size_t)width * height;
    const F3 w = h3(h_w3), h = h3(h_h3);
    // cuda/kernels/gen_rays.cuh:636-646
    const F3 dw = { w.x / width, w.y / width, w.z / width };
    const F3 dh = { h.x / height, h.y / height, h.z / height };
    const F3 dir = h_normalize3(h_cross(w, h));
    plane_parallel_kernel<<<prng_blocks(ctx, n_rays), RAYS_THREADS, 0, (cudaStream_t)stream>>>(
        d_rays, width, n_rays, h3(h_base3), dw, dh, length, dir, seed);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_orthographic_projection_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays,
                                            int resolution_x, int resolution_y,
                                            const float* h_camera_position3, const float* h_look_at3,
                                            const float* h_view_up3, float vertical_extent,
                                            float length, void* stream)
{
    GB_REQUIRE(ctx && d_rays && h_camera_position3 && h_look_at3 && h_view_up3, GRACE_B200_EINVAL,
               "NULL argument");
    GB_REQUIRE(resolution_x > 0 && resolution_y > 0, GRACE_B200_EINVAL, "empty image");
    // cuda/kernels/gen_rays.cuh:687-709
    const size_t n_rays = (size_t)resolution_x * resolution_y;
    const float aspect = (float)resolution_x / resolution_y;
    const float horizontal_extent = vertical_extent * aspect;
    const F3 cam = h3(h_camera_position3), at = h3(h_look_at3);
    F3 dir = { at.x - cam.x, at.y - cam.y, at.z - cam.z };
    dir = h_normalize3(dir);
    F3 v = h_normalize3(h_cross(dir, h3(h_view_up3)));
    F3 u = h_normalize3(h_cross(v, dir));
    v.x = (float)(v.x * (horizontal_extent / 2.)); v.y = (float)(v.y * (horizontal_extent / 2.));
    v.z = (float)(v.z * (horizontal_extent / 2.));
    u.x = (float)(u.x * (vertical_extent / 2.)); u.y = (float)(u.y * (vertical_extent / 2.));
    u.z = (float)(u.z * (vertical_extent / 2.));
    ortho_kernel<<<cap_blocks(ctx, n_rays, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        d_rays, resolution_x, resolution_y, n_rays, cam, dir, v, u, length);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_pinhole_camera_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, int resolution_x,
                                   int resolution_y, const float* h_camera_position3,
                                   const float* h_look_at3, const float* h_view_up3, float fov_y,
                                   float length, void* stream)
{
    GB_REQUIRE(ctx && d_rays && h_camera_position3 && h_look_at3 && h_view_up3, GRACE_B200_EINVAL,
               "NULL argument");
    GB_REQUIRE(resolution_x > 0 && resolution_y > 0, GRACE_B200_EINVAL, "empty image");
    // cuda/kernels/gen_rays.cuh:750-770
    const size_t n_rays = (size_t)resolution_x * resolution_y;
    const float aspect = (float)resolution_x / resolution_y;
    const F3 cam = h3(h_camera_position3), at = h3(h_look_at3);
    const F3 dir = { at.x - cam.x, at.y - cam.y, at.z - cam.z };
    const F3 v = h_normalize3(h_cross(dir, h3(h_view_up3)));
    const F3 u = h_normalize3(h_cross(v, dir));
    F3 n = h_normalize3(dir);
    const float pref = (float)(1. / std::tan(fov_y / 2.));
    n.x *= pref; n.y *= pref; n.z *= pref;
    pinhole_kernel<<<cap_blocks(ctx, n_rays, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        d_rays, resolution_x, resolution_y, n_rays, aspect, cam, v, u, n, length);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_healpix_rays(grace_b200_ctx* ctx, grace_b200_ray* d_rays, size_t n_rays, long nside,
                            long first_pixel, float ox, float oy, float oz, float length, void* stream)
{
    GB_REQUIRE(ctx && (d_rays || n_rays == 0), GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(nside > 0 && (nside & (nside - 1)) == 0 && nside <= 8192, GRACE_B200_EINVAL,
               "nside must be a power of two <= 8192 (NESTED scheme, 32-bit pixel index)");
    GB_REQUIRE(first_pixel >= 0 && first_pixel + (long)n_rays <= 12 * nside * nside, GRACE_B200_EINVAL,
               "pixel range outside [0, 12 nside^2)");
    if (n_rays == 0) return GRACE_B200_OK;
    healpix_kernel<<<cap_blocks(ctx, n_rays, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        d_rays, n_rays, nside, first_pixel, ox, oy, oz, length);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // extern "C"
