// radix_sort.cu -- stable key/index onesweep radix sort (8-bit digits, decoupled
// look-back) plus the record gathers that apply the permutation.
//
// Replaces thrust::sort_by_key(keys, float4 / Ray) in GRACE (cuda/build_sph.cuh:46,57,70,81;
// cuda/kernels/gen_rays.cuh:483,520,577,615).  Same contract: ascending, stable.
//
// B200 design (HBM-bound integer work, no tensor cores):
//   * ONE histogram pass over the keys builds the digit counts of every pass;
//   * each digit pass is a single kernel ("onesweep"): a CTA takes the next tile by
//     ticket, ranks its keys with warp match-any + per-warp digit counters in shared
//     memory, publishes the tile's digit counts, resolves its global offsets by
//     decoupled look-back over earlier tiles, reorders the tile in shared memory and
//     writes each digit run contiguously (coalesced);
//   * only (key, 32-bit index) pairs move through the passes -- the 16 B / 28 B
//     records are gathered once at the end -- so a 4-pass sort of float4 spheres moves
//     ~100 B/particle instead of 4 * 2 * 20 = 160 B.
//   Per pass the algorithmic traffic is 2*(sizeof(Key)+4) B per element.
#include "radix_sort.cuh"

namespace {

constexpr unsigned FLAG_AGG  = 1u << 30;   // tile aggregate available
constexpr unsigned FLAG_INCL = 2u << 30;   // inclusive prefix available
constexpr unsigned FLAG_MASK = 3u << 30;
constexpr unsigned VAL_MASK  = ~FLAG_MASK;

constexpr int HIST_THREADS = 256;
constexpr int HIST_IPT = 8;
constexpr int MAX_PASSES = 8;

template <typename KeyT>
__global__ void __launch_bounds__(HIST_THREADS)
hist_kernel(const KeyT* __restrict__ keys, size_t n, int passes, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t sh[MAX_PASSES * GB_RADIX];
    for (int i = threadIdx.x; i < passes * GB_RADIX; i += HIST_THREADS) sh[i] = 0;
    __syncthreads();
    const size_t chunk = (size_t)HIST_THREADS * HIST_IPT;
    for (size_t base = (size_t)blockIdx.x * chunk; base < n; base += (size_t)gridDim.x * chunk) {
        const bool full = base + chunk <= n;
#pragma unroll
        for (int i = 0; i < HIST_IPT; ++i) {
            const size_t idx = base + (size_t)i * HIST_THREADS + threadIdx.x;
            if (full) {
                const KeyT k = keys[idx];
                for (int p = 0; p < passes; ++p) {
                    const unsigned d = (unsigned)(k >> (GB_RADIX_BITS * p)) & (GB_RADIX - 1);
                    int all_same;
                    __match_all_sync(0xffffffffu, d, &all_same);
                    if (all_same) { if ((threadIdx.x & 31) == 0) atomicAdd(&sh[p * GB_RADIX + d], 32u); }
                    else atomicAdd(&sh[p * GB_RADIX + d], 1u);
                }
            } else if (idx < n) {
                const KeyT k = keys[idx];
                for (int p = 0; p < passes; ++p)
                    atomicAdd(&sh[p * GB_RADIX + ((unsigned)(k >> (GB_RADIX_BITS * p)) & (GB_RADIX - 1))], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * GB_RADIX; i += HIST_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// One block per pass: exclusive scan of the 256 digit counts.
__global__ void __launch_bounds__(GB_RADIX)
digit_scan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ base)
{
    __shared__ uint32_t warp_tot[GB_RADIX / 32];
    const int p = blockIdx.x, d = threadIdx.x;
    const uint32_t c = hist[p * GB_RADIX + d];
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((d & 31) >= o) incl += t;
    }
    if ((d & 31) == 31) warp_tot[d >> 5] = incl;
    __syncthreads();
    uint32_t add = 0;
    for (int w = 0; w < (d >> 5); ++w) add += warp_tot[w];
    base[p * GB_RADIX + d] = add + incl - c;
}

#ifndef OS_MINB
#define OS_MINB 3
#endif
#ifndef OS_NT32
#define OS_NT32 384
#endif
#ifndef OS_IPT32
#define OS_IPT32 16
#endif
#ifndef OS_NT64
#define OS_NT64 384
#endif
#ifndef OS_IPT64
#define OS_IPT64 12
#endif
// GATHER (last pass of the keys + sort entry point): instead of the (key, index) pair, the 16-byte record
// the index names is written to the sorted position -- the separate gather launch read the permutation
// back and wrote the same records; the keys are still written when the caller wants them.
template <typename KeyT, int NT, int IPT, bool GATHER>
__global__ void __launch_bounds__(NT, OS_MINB)
onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                uint32_t n, int shift, const uint32_t* __restrict__ base,
                uint32_t* __restrict__ tile_state, uint32_t* __restrict__ tile_counter,
                const float4* __restrict__ rec_in, float4* __restrict__ rec_out)
{
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * IPT;
    static_assert(NT >= GB_RADIX, "need one thread per digit");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* warp_hist = (uint32_t*)smem_raw;              // [NW][256]
    uint32_t* digit_start = warp_hist + NW * GB_RADIX;      // [256]
    uint32_t* digit_global = digit_start + GB_RADIX;        // [256]
    KeyT* s_keys = (KeyT*)(digit_global + GB_RADIX);        // [TILE]
    uint32_t* s_vals = (uint32_t*)(s_keys + TILE);          // [TILE]
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_tot[GB_RADIX / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < NW * GB_RADIX; i += NT) warp_hist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t n_valid = min((uint32_t)TILE, n - tile_base);
    const uint32_t n_invalid = (uint32_t)TILE - n_valid;

    // ---- load keys, warp-striped: element order = (warp, item, lane) ----
    KeyT key[IPT];
    uint32_t rank[IPT];
    const uint32_t warp_base = tile_base + (uint32_t)warp * (IPT * 32);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t idx = warp_base + i * 32 + lane;
        key[i] = idx < n ? keys_in[idx] : ~KeyT(0);
    }
#ifndef OS_LATE_VALS
    uint32_t val[IPT];     // issued with the keys: their latency hides behind the ranking
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const uint32_t idx = warp_base + i * 32 + lane;
        val[i] = vals_in ? (idx < n ? vals_in[idx] : 0u) : idx;
    }
#endif
    // ---- rank within the warp (stable) ----
    uint32_t* wh = warp_hist + warp * GB_RADIX;
    const unsigned lt = gb_lanemask_lt();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned d = (unsigned)(key[i] >> shift) & (GB_RADIX - 1);
#ifdef OS_MATCH
        const unsigned peers = __match_any_sync(0xffffffffu, d);
#else
        // lanes holding the same digit, from one ballot per digit bit: MATCH.ANY is a
        // long-latency instruction and sixteen dependent ones per thread dominated the pass
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < GB_RADIX_BITS; ++b) {
            const bool bit = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
        }
#endif
        const unsigned lower = __popc(peers & lt);
        uint32_t b = 0;
        if (lower == 0) { b = wh[d]; wh[d] = b + __popc(peers); }
        b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
        rank[i] = b + lower;
        __syncwarp();
    }
    __syncthreads();

    // ---- per-digit: prefix over warps, publish, scan over digits, look-back ----
    uint32_t my_count = 0, my_sum = 0;
    if (tid < GB_RADIX) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t t = warp_hist[w * GB_RADIX + tid];
            warp_hist[w * GB_RADIX + tid] = sum;
            sum += t;
        }
        my_sum = sum;
        my_count = sum - (tid == GB_RADIX - 1 ? n_invalid : 0u);
        gb_st_volatile_u32(tile_state + (size_t)tile * GB_RADIX + tid,
                           my_count | (tile == 0 ? FLAG_INCL : FLAG_AGG));
        // exclusive scan of my_sum across the 256 digit threads
        uint32_t incl = my_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        my_sum = incl - my_sum;     // exclusive within warp
    }
    __syncthreads();
    if (tid < GB_RADIX) {
        uint32_t add = 0;
        for (int w = 0; w < warp; ++w) add += s_warp_tot[w];
        const uint32_t start = my_sum + add;
        digit_start[tid] = start;
        // decoupled look-back over earlier tiles
        uint32_t excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            for (;;) {
                const uint32_t v = gb_ld_volatile_u32(tile_state + (size_t)t * GB_RADIX + tid);
                const uint32_t f = v & FLAG_MASK;
                if (f == 0) continue;            // predecessor not published yet: spin
                excl += v & VAL_MASK;
                if (f == FLAG_INCL) break;
                --t;
            }
            gb_st_volatile_u32(tile_state + (size_t)tile * GB_RADIX + tid,
                               (excl + my_count) | FLAG_INCL);
        }
        digit_global[tid] = base[tid] + excl - start;
    }
    __syncthreads();

    // ---- reorder the tile in shared memory ----
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const unsigned d = (unsigned)(key[i] >> shift) & (GB_RADIX - 1);
        const uint32_t pos = digit_start[d] + wh[d] + rank[i];
        const uint32_t idx = warp_base + i * 32 + lane;
        s_keys[pos] = key[i];
#ifndef OS_LATE_VALS
        s_vals[pos] = val[i]; (void)idx;
#else
        s_vals[pos] = vals_in ? (idx < n ? vals_in[idx] : 0u) : idx;
#endif
    }
    __syncthreads();
    // ---- write digit runs contiguously ----
    if (GATHER) {
#pragma unroll
        for (int i0 = 0; i0 < IPT; i0 += 4) {
            float4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t p = (uint32_t)tid + (uint32_t)(i0 + u) * NT;
                if (p < n_valid) r[u] = __ldg(rec_in + s_vals[p]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t p = (uint32_t)tid + (uint32_t)(i0 + u) * NT;
                if (p < n_valid) {
                    const KeyT k = s_keys[p];
                    const unsigned d = (unsigned)(k >> shift) & (GB_RADIX - 1);
                    const uint32_t dst = digit_global[d] + p;
                    rec_out[dst] = r[u];
                    if (keys_out) keys_out[dst] = k;
                }
            }
        }
    } else {
        for (uint32_t p = tid; p < n_valid; p += NT) {
            const KeyT k = s_keys[p];
            const unsigned d = (unsigned)(k >> shift) & (GB_RADIX - 1);
            const uint32_t dst = digit_global[d] + p;
            keys_out[dst] = k;
            vals_out[dst] = s_vals[p];
        }
    }
}

template <typename KeyT> struct SortCfg;
template <> struct SortCfg<uint32_t> { static constexpr int NT = OS_NT32, IPT = OS_IPT32; };
template <> struct SortCfg<uint64_t> { static constexpr int NT = OS_NT64, IPT = OS_IPT64; };

template <typename KeyT>
constexpr size_t onesweep_smem()
{
    using C = SortCfg<KeyT>;
    return (size_t)(C::NT / 32) * GB_RADIX * 4 + 2 * GB_RADIX * 4 +
           (size_t)C::NT * C::IPT * (sizeof(KeyT) + 4);
}

inline size_t n_tiles_for(size_t n, int tile) { return (n + tile - 1) / tile; }

// ---- gathers --------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather16_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                const uint32_t* __restrict__ perm, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __ldg(in + perm[i]);
}

__global__ void __launch_bounds__(256)
gather4_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
               const uint32_t* __restrict__ perm, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __ldg(in + perm[i]);
}

// 28-byte records (grace::Ray): one thread per 4-byte word so that both the
// source record and the destination are touched with unit stride.
__global__ void __launch_bounds__(256)
gather28_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                const uint32_t* __restrict__ perm, size_t n)
{
    const size_t total = n * 7;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const size_t i = t / 7;
        const unsigned c = (unsigned)(t - i * 7);
        out[t] = __ldg(in + (size_t)perm[i] * 7 + c);
    }
}

} // namespace

size_t gb_sort_workspace_bytes(size_t n, int key_bytes)
{
    const int tile = key_bytes == 8 ? SortCfg<uint64_t>::NT * SortCfg<uint64_t>::IPT
                                    : SortCfg<uint32_t>::NT * SortCfg<uint32_t>::IPT;
    const size_t tiles = n_tiles_for(n, tile);
    size_t b = 0;
    b += 2 * gb_align(n * (size_t)key_bytes);            // key ping/pong
    b += 2 * gb_align(n * 4);                            // index ping/pong
    b += gb_align((size_t)MAX_PASSES * GB_RADIX * 4);    // hist
    b += gb_align((size_t)MAX_PASSES * GB_RADIX * 4);    // base
    b += gb_align((size_t)MAX_PASSES * 4);               // tile counters
    b += gb_align((size_t)MAX_PASSES * tiles * GB_RADIX * 4);  // tile states
    return b + 256;
}

template <typename KeyT>
int gb_sort_pairs(grace_b200_ctx* ctx, const KeyT* d_keys_in, KeyT* d_keys_out,
                  uint32_t* d_perm_out, size_t n, int key_bits, void* ws,
                  const uint32_t* d_hist_in, cudaStream_t st, const void* d_rec16_in, void* d_rec16_out)
{
    using C = SortCfg<KeyT>;
    constexpr int TILE = C::NT * C::IPT;
    GB_REQUIRE(n < (1ull << 30), GRACE_B200_ERANGE, "sort of %zu elements exceeds 2^30", n);
    GB_REQUIRE(key_bits > 0 && key_bits <= (int)sizeof(KeyT) * 8, GRACE_B200_EINVAL, "bad key_bits");
    if (n == 0) return GRACE_B200_OK;
    const int passes = (key_bits + GB_RADIX_BITS - 1) / GB_RADIX_BITS;
    const size_t tiles = n_tiles_for(n, TILE);

    GbArena a(ws, gb_sort_workspace_bytes(n, (int)sizeof(KeyT)));
    KeyT* kbuf[2] = { a.take<KeyT>(n), a.take<KeyT>(n) };
    uint32_t* vbuf[2] = { a.take<uint32_t>(n), a.take<uint32_t>(n) };
    uint32_t* hist = a.take<uint32_t>((size_t)MAX_PASSES * GB_RADIX);
    uint32_t* base = a.take<uint32_t>((size_t)MAX_PASSES * GB_RADIX);
    uint32_t* counters = a.take<uint32_t>(MAX_PASSES);
    uint32_t* states = a.take<uint32_t>((size_t)MAX_PASSES * tiles * GB_RADIX);

    // hist | base | counters | states are contiguous: one memset clears what must be zero.
    if (d_hist_in) {
        GB_CUDA(cudaMemsetAsync(counters, 0, (char*)(states + (size_t)passes * tiles * GB_RADIX) - (char*)counters, st));
    } else {
        GB_CUDA(cudaMemsetAsync(hist, 0, (char*)(states + (size_t)passes * tiles * GB_RADIX) - (char*)hist, st));
        const size_t chunk = (size_t)HIST_THREADS * HIST_IPT;
        size_t blocks = (n + chunk - 1) / chunk;
        const size_t cap = (size_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        hist_kernel<KeyT><<<(int)blocks, HIST_THREADS, 0, st>>>(d_keys_in, n, passes, hist);
        GB_LAUNCH_CHECK();
    }
    digit_scan_kernel<<<passes, GB_RADIX, 0, st>>>(d_hist_in ? d_hist_in : hist, base);
    GB_LAUNCH_CHECK();

    constexpr size_t smem = onesweep_smem<KeyT>();
    static_assert(C::IPT % 4 == 0, "the gathering write phase handles four items at a time");
    GB_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, C::NT, C::IPT, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT, C::NT, C::IPT, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool fuse = d_rec16_in && d_rec16_out && passes > 1;
    const KeyT* kin = d_keys_in;
    const uint32_t* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
        const bool last = (p == passes - 1);
        KeyT* kout = (last && passes > 1) ? d_keys_out : kbuf[p & 1];
        uint32_t* vout = (last && passes > 1) ? d_perm_out : vbuf[p & 1];
        if (last && fuse)       // records to their sorted places; keys only if asked for, the permutation not at all
            onesweep_kernel<KeyT, C::NT, C::IPT, true><<<(int)tiles, C::NT, smem, st>>>(
                kin, vin, d_keys_out, nullptr, (uint32_t)n, p * GB_RADIX_BITS, base + p * GB_RADIX,
                states + (size_t)p * tiles * GB_RADIX, counters + p, (const float4*)d_rec16_in, (float4*)d_rec16_out);
        else
            onesweep_kernel<KeyT, C::NT, C::IPT, false><<<(int)tiles, C::NT, smem, st>>>(
                kin, vin, kout, vout, (uint32_t)n, p * GB_RADIX_BITS, base + p * GB_RADIX,
                states + (size_t)p * tiles * GB_RADIX, counters + p, nullptr, nullptr);
        GB_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
    }
    if (passes == 1) {  // single pass cannot write over its own input
        GB_CUDA(cudaMemcpyAsync(d_keys_out, kbuf[0], n * sizeof(KeyT), cudaMemcpyDeviceToDevice, st));
        GB_CUDA(cudaMemcpyAsync(d_perm_out, vbuf[0], n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return GRACE_B200_OK;
}

template int gb_sort_pairs<uint32_t>(grace_b200_ctx*, const uint32_t*, uint32_t*, uint32_t*, size_t,
                                     int, void*, const uint32_t*, cudaStream_t, const void*, void*);
template int gb_sort_pairs<uint64_t>(grace_b200_ctx*, const uint64_t*, uint64_t*, uint32_t*, size_t,
                                     int, void*, const uint32_t*, cudaStream_t, const void*, void*);

int gb_gather_records(const void* d_in, void* d_out, const uint32_t* d_perm, size_t n,
                      int rec_bytes, int sm_count, cudaStream_t st)
{
    if (n == 0) return GRACE_B200_OK;
    const size_t work = rec_bytes == 28 ? n * 7 : n;
    size_t blocks = (work + 255) / 256;
    const size_t cap = (size_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (rec_bytes == 16)
        gather16_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)d_in, (float4*)d_out, d_perm, n);
    else if (rec_bytes == 4)
        gather4_kernel<<<(int)blocks, 256, 0, st>>>((const uint32_t*)d_in, (uint32_t*)d_out, d_perm, n);
    else if (rec_bytes == 28)
        gather28_kernel<<<(int)blocks, 256, 0, st>>>((const uint32_t*)d_in, (uint32_t*)d_out, d_perm, n);
    else
        return gb_set_error(GRACE_B200_EINVAL, "unsupported record size %d (4, 16 or 28)", rec_bytes);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

namespace {

template <typename KeyT>
int sort_pairs_api(grace_b200_ctx* ctx, KeyT* d_keys, void* d_values, int value_bytes, size_t n,
                   int key_bits, uint32_t* d_perm, void* stream)
{
    GB_REQUIRE(ctx && d_keys, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(d_values == nullptr || value_bytes == 4 || value_bytes == 16 || value_bytes == 28,
               GRACE_B200_EINVAL, "value_bytes must be 4, 16 or 28");
    if (n == 0) return GRACE_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sort_ws = gb_sort_workspace_bytes(n, (int)sizeof(KeyT));
    const size_t perm_bytes = gb_align(n * 4);
    const size_t val_bytes = d_values ? gb_align(n * (size_t)value_bytes) : 0;
    char* ws = (char*)gb_workspace(ctx, sort_ws + perm_bytes + val_bytes);
    if (!ws) return GRACE_B200_ENOMEM;
    uint32_t* perm = (uint32_t*)(ws + sort_ws);
    void* vtmp = ws + sort_ws + perm_bytes;
    int rc = gb_sort_pairs<KeyT>(ctx, d_keys, d_keys, perm, n, key_bits, ws, nullptr, st);
    if (rc) return rc;
    if (d_values) {
        rc = gb_gather_records(d_values, vtmp, perm, n, value_bytes, ctx->sm_count, st);
        if (rc) return rc;
        GB_CUDA(cudaMemcpyAsync(d_values, vtmp, n * (size_t)value_bytes, cudaMemcpyDeviceToDevice, st));
    }
    if (d_perm) GB_CUDA(cudaMemcpyAsync(d_perm, perm, n * 4, cudaMemcpyDeviceToDevice, st));
    return GRACE_B200_OK;
}

} // namespace

extern "C" {

int grace_b200_sort_pairs_u32(grace_b200_ctx* ctx, uint32_t* d_keys, void* d_values,
                              int value_bytes, size_t n, int key_bits, uint32_t* d_perm,
                              void* stream)
{
    return sort_pairs_api<uint32_t>(ctx, d_keys, d_values, value_bytes, n, key_bits, d_perm, stream);
}

int grace_b200_sort_pairs_u64(grace_b200_ctx* ctx, uint64_t* d_keys, void* d_values,
                              int value_bytes, size_t n, int key_bits, uint32_t* d_perm,
                              void* stream)
{
    return sort_pairs_api<uint64_t>(ctx, d_keys, d_values, value_bytes, n, key_bits, d_perm, stream);
}

} // extern "C"
