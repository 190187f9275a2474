// hits.cu -- per-ray hit lists: count -> exclusive scan -> fill, and the per-ray
// stable sort of hits by distance.
//
// Reference behaviour (GRACE): trace_sph / trace_with_sentinels_sph,
// cuda/trace_sph.cuh:112-241 (hit-count pass, two blocking element read-backs,
// thrust::exclusive_scan, three resizes, fill pass) and sort_by_distance,
// cuda/sort.cuh:100-131 (new sgpu context per call, thrust::sequence,
// sgpu::SegSortPairsFromIndices = blocksort + log2(tiles) global merge passes over ALL
// hits, then two copy+gather passes).
//
// B200 design:
//   * one single-pass decoupled look-back scan turns counts into offsets and a 64-bit
//     total (one read-back instead of two);
//   * the segmented sort never merges across segments: ray segments are binned by
//     length (capacities 32, 512, 1024, ... 8192) and each is sorted entirely in shared
//     memory by one CTA (stable LSD radix on the distance bits, see segsort_smem_kernel);
//     the permutation is applied to indices and payload in the same kernel, so each hit
//     is read and written once: 24 B/hit of HBM traffic against
//     ~12 B x (2 + 2*log2(tiles)) + 32 B for the reference.  Segments longer than the
//     largest capacity take a global-memory bitonic path.
#include "common.cuh"

namespace {

constexpr unsigned long long ST_AGG = 1ull << 62, ST_INCL = 2ull << 62, ST_MASK = 3ull << 62;
constexpr int SC_THREADS = 256, SC_IPT = 8, SC_TILE = SC_THREADS * SC_IPT;

// Exclusive scan of int32 with 64-bit running total.  If add_index != 0, out[i] also
// gets + i (trace_with_sentinels_sph, trace_sph.cuh:196-207).
__global__ void __launch_bounds__(SC_THREADS)
scan_kernel(const int* __restrict__ in, int* __restrict__ out, size_t n, int add_index,
            unsigned long long* __restrict__ block_state, unsigned* __restrict__ ticket,
            long long* __restrict__ total_out)
{
    __shared__ unsigned s_bid;
    __shared__ long long s_warp_tot[SC_THREADS / 32];
    __shared__ long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const size_t base = (size_t)bid * SC_TILE + (size_t)tid * SC_IPT;
    int v[SC_IPT];
    long long tsum = 0;
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        tsum += v[i];
    }
    long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    long long add = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w < warp) add += s_warp_tot[w];
        block_total += s_warp_tot[w];
    }
    if (tid == 0) {
        unsigned long long excl = 0;
        if (bid == 0) {
            gb_st_volatile_u64(block_state, (unsigned long long)block_total | ST_INCL);
        } else {
            gb_st_volatile_u64(block_state + bid, (unsigned long long)block_total | ST_AGG);
            int t = (int)bid - 1;
            for (;;) {
                const unsigned long long s = gb_ld_volatile_u64(block_state + t);
                const unsigned long long f = s & ST_MASK;
                if (f == 0) continue;
                excl += s & ~ST_MASK;
                if (f == ST_INCL) break;
                --t;
            }
            gb_st_volatile_u64(block_state + bid, (excl + (unsigned long long)block_total) | ST_INCL);
        }
        s_excl = (long long)excl;
        if ((size_t)bid == (n - 1) / SC_TILE && total_out)
            *total_out = (long long)excl + block_total + (add_index ? (long long)n : 0);
    }
    __syncthreads();
    long long run = s_excl + add + incl - tsum;
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i) {
        if (base + i < n) out[base + i] = (int)(run + (add_index ? (long long)(base + i) : 0));
        run += v[i];
    }
}

int run_scan(grace_b200_ctx* ctx, const int* d_in, int* d_out, size_t n, int add_index,
             long long* d_total, cudaStream_t st)
{
    if (n == 0) {
        if (d_total) GB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(long long), st));
        return GRACE_B200_OK;
    }
    const size_t blocks = (n + SC_TILE - 1) / SC_TILE;
    unsigned long long* state = (unsigned long long*)gb_workspace(ctx, blocks * 8 + 256);
    if (!state) return GRACE_B200_ENOMEM;
    unsigned* ticket = (unsigned*)(ctx->d_scalars + GB_SC_TICKET1);
    GB_CUDA(cudaMemsetAsync(state, 0, blocks * 8, st));
    GB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    scan_kernel<<<(int)blocks, SC_THREADS, 0, st>>>(d_in, d_out, n, add_index, state, ticket, d_total);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

// ---------------------------------------------------------------------------
// segmented sort
// ---------------------------------------------------------------------------
// Float distance -> unsigned key with the same ordering as operator< on floats
// (-0 and +0 compare equal, so both map to the key of +0).
__device__ __forceinline__ unsigned dist_key(float f)
{
    unsigned b = __float_as_uint(f);
    if ((b << 1) == 0u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

constexpr int N_CLASSES = 10;
constexpr int N_CLASS_SLOTS = 16;        // counter slots reserved in d_scalars (GB_SC_CLASS)
// Class capacities (elements); the last class = anything larger (global path).  A CTA always
// sorts CAP slots (the tail is padding), so closely spaced capacities keep most of every CTA's
// work useful: orthographic tiles through 2^24 particles have 2049-4869 hits per ray, which a
// single 8192 class sorted at 28 % efficiency (5.7 of the 7.6 ms per tile).
__host__ __device__ constexpr int class_cap(int c)
{
    return c == 0 ? 32 : c == 1 ? 512 : c == 2 ? 1024 : c == 3 ? 1536 : c == 4 ? 2048 : c == 5 ? 3072
         : c == 6 ? 4096 : c == 7 ? 6144 : c == 8 ? 8192 : 0x7fffffff;
}

// Bin segments by length.  lists: [N_CLASSES][n_rays] ray ids; counts: [N_CLASSES];
// xl_total: total elements in last-class segments (padded to pow2 per segment).
__global__ void __launch_bounds__(256)
classify_kernel(const int* __restrict__ offsets, int n_rays, long long total,
                int* __restrict__ lists, int* __restrict__ counts,
                unsigned long long* __restrict__ xl_total, unsigned long long* __restrict__ xl_offsets)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const long long b = offsets[r];
    const long long e = (r + 1 < n_rays) ? (long long)offsets[r + 1] : total;
    const long long len = e - b;
    if (len <= 1) return;
    int c = 0;
    while (len > class_cap(c)) ++c;
    const int slot = atomicAdd(counts + c, 1);
    lists[(size_t)c * n_rays + slot] = r;
    if (c == N_CLASSES - 1) {
        unsigned long long m = 1;
        while ((long long)m < len) m <<= 1;
        xl_offsets[slot] = atomicAdd(xl_total, m);
    }
}

// One CTA sorts one segment entirely in shared memory: a stable LSD radix sort (four 8-bit passes)
// on (distance key, local position) pairs.  Elements are owned (warp, item, lane) -> ascending
// position, lanes holding the same digit are found with one ballot per digit bit, every warp keeps
// its own digit counters, and a digit-major / warp-minor scan gives each warp its base per digit:
// 3 barriers per pass where the bitonic network it replaces needed one per compare-exchange stage
// (66 for 2048 elements, 91 for 8192).  A pass whose digit is the same for the whole segment is
// skipped.  The permutation is then applied to distances, indices and payload in place.
// (capped at 56 registers: the kernel is a chain of barrier-separated phases and more resident CTAs hide them;
// 96-128 registers -> 5.31 ms, 80 -> 4.78, 64 -> 4.75, 56 -> 4.61, 48 -> 4.62 for the 1.68e8 hits of 2^17 rays)
constexpr int SS_REGS = 56;
template <int CAP, int NT>
__global__ void __launch_bounds__(NT, 65536 / (NT * SS_REGS) > 16 ? 16 : 65536 / (NT * SS_REGS))
segsort_smem_kernel(float* __restrict__ dist, const int* __restrict__ offsets, int n_rays,
                    long long total, int* __restrict__ idx, unsigned* __restrict__ data,
                    const int* __restrict__ list, const int* __restrict__ count_ptr)
{
    constexpr int NW = NT / 32;
    constexpr int IPT = CAP / NT;
    static_assert(CAP % NT == 0 && CAP <= 65536, "positions are 16-bit");
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned* k_buf[2] = { (unsigned*)ss_smem, (unsigned*)ss_smem + CAP };
    unsigned short* p_buf[2] = { (unsigned short*)(ss_smem + 8 * CAP), (unsigned short*)(ss_smem + 8 * CAP) + CAP };
    unsigned short* wh_all = (unsigned short*)(ss_smem + 12 * CAP);       // [NW][256] per-warp digit counters / bases
    unsigned* dig_tot = (unsigned*)(ss_smem + 12 * CAP + NW * 256 * 2);   // [256] per-digit totals, then bases
    __shared__ unsigned s_skip;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = gb_lanemask_lt();
    unsigned short* wh = wh_all + warp * 256;
    const int n_seg = *count_ptr;
    for (int s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const int r = list[s];
        const long long b = offsets[r];
        const long long e = (r + 1 < n_rays) ? (long long)offsets[r + 1] : total;
        const int len = (int)(e - b);
        int cur = 0;
        // element order = (warp, item, lane): position q = warp * IPT * 32 + i * 32 + lane
        for (int i = 0; i < IPT; ++i) {
            const int q = (warp * IPT + i) * 32 + lane;
            k_buf[0][q] = q < len ? dist_key(dist[b + q]) : 0xffffffffu;     // padding sorts last, stably
            p_buf[0][q] = (unsigned short)q;
        }
        __syncthreads();
        for (int shift = 0; shift < 32; shift += 8) {
            for (int i = tid; i < NW * 256; i += NT) wh_all[i] = 0;
            __syncthreads();
            unsigned key[IPT], rank[IPT];
            unsigned short pos[IPT];
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const int q = (warp * IPT + i) * 32 + lane;
                key[i] = k_buf[cur][q];
                pos[i] = p_buf[cur][q];
                const unsigned d = (key[i] >> shift) & 255u;
                unsigned peers = 0xffffffffu;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const bool one = (d >> bit) & 1u;
                    const unsigned m = __ballot_sync(0xffffffffu, one);
                    peers &= one ? m : ~m;
                }
                const unsigned lower = __popc(peers & lt);
                unsigned base = 0;
                if (lower == 0) { base = wh[d]; wh[d] = (unsigned short)(base + __popc(peers)); }
                base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
                rank[i] = base + lower;
                __syncwarp();
            }
            __syncthreads();
            // digit-major, warp-minor exclusive scan: a thread owns digit d
            for (int d = tid; d < 256; d += NT) {
                unsigned run = 0;
                for (int w = 0; w < NW; ++w) {
                    const unsigned c = wh_all[w * 256 + d];
                    wh_all[w * 256 + d] = (unsigned short)run;       // this warp's base inside the digit
                    run += c;
                }
                dig_tot[d] = run;
            }
            __syncthreads();
            if (warp == 0) {       // 256 totals -> exclusive bases, 8 per lane
                unsigned loc[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { loc[k] = dig_tot[lane * 8 + k]; sum += loc[k]; }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                unsigned run = incl - sum;
#pragma unroll
                for (int k = 0; k < 8; ++k) { dig_tot[lane * 8 + k] = run; run += loc[k]; }
                // every valid element in one digit (padding keys are all ones: digit 255): nothing moves
                bool full = false;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    full |= loc[k] == (unsigned)CAP || (lane * 8 + k != 255 && loc[k] == (unsigned)len);
                const bool any_full = __any_sync(0xffffffffu, full);
                if (lane == 0) s_skip = any_full ? 1u : 0u;
            }
            __syncthreads();
            const bool skip = s_skip != 0u;
            if (!skip) {
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const unsigned d = (key[i] >> shift) & 255u;
                    const unsigned dst = dig_tot[d] + wh[d] + rank[i];
                    k_buf[cur ^ 1][dst] = key[i];
                    p_buf[cur ^ 1][dst] = pos[i];
                }
                cur ^= 1;
            }
            __syncthreads();
        }
        // apply the permutation: read everything first, then write (in place)
        float dv[IPT]; int iv[IPT]; unsigned pv[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = tid + q * NT;
            if (i < len) {
                const unsigned src = p_buf[cur][i];
                dv[q] = dist[b + src]; iv[q] = idx[b + src]; pv[q] = data[b + src];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = tid + q * NT;
            if (i < len) { dist[b + i] = dv[q]; idx[b + i] = iv[q]; data[b + i] = pv[q]; }
        }
        __syncthreads();
    }
}

// Global-memory path for segments above the shared-memory capacity: one CTA per
// segment, bitonic network on a padded power-of-two scratch range.
__global__ void __launch_bounds__(1024)
segsort_global_kernel(float* __restrict__ dist, const int* __restrict__ offsets, int n_rays,
                      long long total, int* __restrict__ idx, unsigned* __restrict__ data,
                      const int* __restrict__ list, const int* __restrict__ count_ptr,
                      const unsigned long long* __restrict__ xl_offsets,
                      unsigned long long* __restrict__ comp_all, float* __restrict__ tmp_d,
                      int* __restrict__ tmp_i, unsigned* __restrict__ tmp_p)
{
    const int n_seg = *count_ptr;
    const int NT = blockDim.x;
    for (int s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const int r = list[s];
        const long long b = offsets[r];
        const long long e = (r + 1 < n_rays) ? (long long)offsets[r + 1] : total;
        const long long len = e - b;
        long long m = 1;
        while (m < len) m <<= 1;
        unsigned long long* comp = comp_all + xl_offsets[s];
        for (long long i = threadIdx.x; i < m; i += NT)
            comp[i] = i < len ? ((unsigned long long)dist_key(dist[b + i]) << 32) | (unsigned)i
                              : ~0ull;
        __syncthreads();
        for (long long k = 2; k <= m; k <<= 1) {
            for (long long j = k >> 1; j > 0; j >>= 1) {
                for (long long t = threadIdx.x; t < (m >> 1); t += NT) {
                    const long long lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const long long hi = lo | j;
                    const bool up = (lo & k) == 0;
                    const unsigned long long a = comp[lo], c = comp[hi];
                    if ((a > c) == up) { comp[lo] = c; comp[hi] = a; }
                }
                __syncthreads();
            }
        }
        float* td = tmp_d + xl_offsets[s]; int* ti = tmp_i + xl_offsets[s]; unsigned* tp = tmp_p + xl_offsets[s];
        for (long long i = threadIdx.x; i < len; i += NT) {
            const unsigned src = (unsigned)(comp[i] & 0xffffffffu);
            td[i] = dist[b + src]; ti[i] = idx[b + src]; tp[i] = data[b + src];
        }
        __syncthreads();
        for (long long i = threadIdx.x; i < len; i += NT) {
            dist[b + i] = td[i]; idx[b + i] = ti[i]; data[b + i] = tp[i];
        }
        __syncthreads();
    }
}

template <int CAP, int NT>
int launch_smem_class(grace_b200_ctx* ctx, int cls, float* dist, const int* offsets, int n_rays,
                      long long total, int* idx, unsigned* data, const int* lists, const int* counts,
                      int h_count, cudaStream_t st)
{
    if (h_count == 0) return GRACE_B200_OK;
    const size_t smem = (size_t)CAP * 12 + (size_t)(NT / 32) * 256 * 2 + 256 * 4;
    GB_CUDA(cudaFuncSetAttribute(segsort_smem_kernel<CAP, NT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, segsort_smem_kernel<CAP, NT>, NT, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = ctx->sm_count * per_sm;
    if (blocks > h_count) blocks = h_count;
    segsort_smem_kernel<CAP, NT><<<blocks, NT, smem, st>>>(
        dist, offsets, n_rays, total, idx, data, lists + (size_t)cls * n_rays, counts + cls);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // namespace

extern "C" {

int grace_b200_exclusive_scan_i32(grace_b200_ctx* ctx, const int* d_in, int* d_out, size_t n,
                                  long long* d_total, void* stream)
{
    GB_REQUIRE(ctx && (n == 0 || (d_in && d_out)), GRACE_B200_EINVAL, "NULL argument");
    return run_scan(ctx, d_in, d_out, n, 0, d_total, (cudaStream_t)stream);
}

// Pass 1 of trace_sph in two halves, so that a caller can put other work on the GPU between them:
// enqueue = hit counts + exclusive scan + asynchronous read-back of the total and of the traversal's
// error flag; finish = wait for them and check.
static int hits_count_enqueue(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays, const float* d_spheres4,
                              size_t n, const grace_b200_tree* tree, int with_sentinels, int* d_ray_offsets, cudaStream_t st)
{
    // one traversal that counts AND records the hits (trace.cu) where it applies, else the plain counting traversal
    int rc = gb_trace_record_f4(ctx, d_rays, n_rays, d_spheres4, n, tree, d_ray_offsets, st);
    if (rc) rc = grace_b200_trace_hitcounts_f4(ctx, d_rays, n_rays, d_spheres4, n, tree, d_ray_offsets, (void*)st);
    if (rc) return rc;
    long long* d_total = (long long*)(ctx->d_scalars + GB_SC_TOTAL64);
    rc = run_scan(ctx, d_ray_offsets, d_ray_offsets, n_rays, with_sentinels ? 1 : 0, d_total, st);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_TOTAL64, d_total, sizeof(long long), cudaMemcpyDeviceToHost, st));
    // the traversal's error flag rides on the same synchronisation: counts from a walk that overflowed
    // its stack or did not terminate are short, and the fill pass would disagree with them
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_ERRFLAG, ctx->d_scalars + GB_SC_ERRFLAG, sizeof(int),
                            cudaMemcpyDeviceToHost, st));
    if (ctx->rec_valid) {     // where the stolen subtrees' hits start; did the hit pool of the recording run dry?
        if ((rc = gb_trace_resolve_recorded(ctx, d_ray_offsets, st))) return rc;
        GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_LB, ctx->d_scalars + GB_SC_LB, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        ctx->rec_units = n_rays / 32;
    }
    return GRACE_B200_OK;
}

static int hits_count_finish(grace_b200_ctx* ctx, long long* h_total_hits, cudaStream_t st)
{
    GB_CUDA(cudaStreamSynchronize(st));
    *h_total_hits = *(long long*)(ctx->h_pinned + GB_SC_TOTAL64);
    if (ctx->rec_valid && ctx->h_pinned[GB_SC_LB + 7]) {       // counts are right, the recorded lists are not: two passes
        ctx->rec_valid = 0;
        // ... and the next recording gets the pool this one would have needed: 16 bytes a hit + a partly filled chunk per unit
        const size_t units = ctx->rec_units + (size_t)std::max(ctx->h_pinned[GB_SC_LB + 1], 0);
        const size_t need = (size_t)*h_total_hits * 16 + units * 8192 + ((size_t)64 << 20);
        ctx->rec_pool_learned = std::min<size_t>(std::max(ctx->rec_pool_learned, need + need / 8), (size_t)64 << 30);
    }
    GB_REQUIRE(ctx->h_pinned[GB_SC_ERRFLAG] == 0, GRACE_B200_EDEVICE,
               "device-side traversal error %d (1 = stack overflow, 2 = walk did not terminate): hit counts are incomplete",
               ctx->h_pinned[GB_SC_ERRFLAG]);
    return GRACE_B200_OK;
}

int grace_b200_trace_hits_count_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                   const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                   int with_sentinels, int* d_ray_offsets, long long* h_total_hits,
                                   void* stream)
{
    GB_REQUIRE(ctx && (d_ray_offsets || n_rays == 0) && h_total_hits, GRACE_B200_EINVAL, "NULL argument");
    if (n_rays == 0) { *h_total_hits = 0; return GRACE_B200_OK; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = hits_count_enqueue(ctx, d_rays, n_rays, d_spheres4, n, tree, with_sentinels, d_ray_offsets, st);
    if (rc) return rc;
    if ((rc = hits_count_finish(ctx, h_total_hits, st))) return rc;
    // trace_sph.cuh:117,137: offsets and totals are int in the reference
    GB_REQUIRE(*h_total_hits <= 0x7fffffffLL, GRACE_B200_ERANGE,
               "%lld hits exceed the 32-bit offsets of the reference layout; tile the rays",
               *h_total_hits);
    return GRACE_B200_OK;
}

// ---------------------------------------------------------------------------
// Sorted hit lists of an arbitrarily large ray set, streamed in ray tiles (a 4096^2
// projection of 2^24 particles has ~5e10 hits, 600 GB -- the reference's one-call trace_sph cannot
// hold it, cuda/trace_sph.cuh:117).  Per tile: count -> scan -> fill -> sort_by_distance -> consumer,
// in buffers sized by a hit budget and reused for every tile; a tile that would exceed the budget is
// halved.  Two buffer sets and two streams: the sort and the consumer of tile k run (on a helper
// context's stream) while the counting traversal of tile k + 1 runs on the caller's stream.
// ---------------------------------------------------------------------------
struct TileBufs { int* offsets; int* idx; float* integ; float* dist; cudaEvent_t filled, consumed; };

static int tiles_reserve(grace_b200_ctx* ctx, size_t hit_budget, size_t tile_rays)
{
    const size_t per_set = gb_align(tile_rays * 4) + 3 * gb_align(hit_budget * 4);
    if (ctx->tile_bytes >= 2 * per_set && ctx->tile_mem) return GRACE_B200_OK;
    if (ctx->tile_mem) { GB_CUDA(cudaDeviceSynchronize()); GB_CUDA(cudaFree(ctx->tile_mem)); ctx->tile_mem = nullptr; ctx->tile_bytes = 0; }
    cudaError_t e = cudaMalloc((void**)&ctx->tile_mem, 2 * per_set);
    if (e != cudaSuccess) return gb_set_error(GRACE_B200_ENOMEM, "hit-list tile buffers: cudaMalloc(%zu) failed: %s", 2 * per_set, cudaGetErrorString(e));
    ctx->tile_bytes = 2 * per_set;
    return GRACE_B200_OK;
}

int grace_b200_trace_sorted_tiles_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays,
                                     const float* d_spheres4, size_t n, const grace_b200_tree* tree,
                                     size_t hit_budget, size_t rays_per_tile, grace_b200_hits_tile_fn consume,
                                     void* user, long long* h_total_hits, void* stream)
{
    GB_REQUIRE(ctx && (d_rays || n_rays == 0) && d_spheres4 && tree, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_rays % 32 == 0, GRACE_B200_EINVAL, "Number of rays must be a multiple of the warp size (32).");
    GB_REQUIRE(hit_budget >= 32 && hit_budget <= 0x7fffffffull, GRACE_B200_EINVAL, "hit budget must be in [32, 2^31)");
    if (h_total_hits) *h_total_hits = 0;
    if (n_rays == 0) return GRACE_B200_OK;
    cudaStream_t sa = (cudaStream_t)stream;
    size_t tile = rays_per_tile ? (rays_per_tile + 31) / 32 * 32 : 65536;
    if (tile > n_rays) tile = n_rays;
    // helper context + stream for the sort / consumer side (a context serves one stream at a time)
    if (!ctx->aux) {
        int rc = grace_b200_create(&ctx->aux, ctx->device);
        if (rc) return rc;
        GB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    }
    cudaStream_t sb = ctx->aux_stream;
    int rc = tiles_reserve(ctx, hit_budget, tile);
    if (rc) return rc;
    // the counting traversal of a tile also records its hits (16 bytes each, blocks padded): room for a full tile
    ctx->rec_pool_hint = hit_budget * 24 + ((size_t)64 << 20);
    TileBufs B[2];
    {
        const size_t per_set = ctx->tile_bytes / 2;
        for (int k = 0; k < 2; ++k) {
            char* p = ctx->tile_mem + k * per_set;
            B[k].offsets = (int*)p; p += gb_align(tile * 4);
            B[k].idx = (int*)p; p += gb_align(hit_budget * 4);
            B[k].integ = (float*)p; p += gb_align(hit_budget * 4);
            B[k].dist = (float*)p;
            GB_CUDA(cudaEventCreateWithFlags(&B[k].filled, cudaEventDisableTiming));
            GB_CUDA(cudaEventCreateWithFlags(&B[k].consumed, cudaEventDisableTiming));
        }
    }
    auto cleanup = [&](int code) {
        ctx->rec_pool_hint = 0;
        cudaStreamSynchronize(sb);
        cudaStreamSynchronize(sa);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(B[k].filled); cudaEventDestroy(B[k].consumed); }
        return code;
    };
    long long grand = 0;
    size_t first = 0;           // first ray of the tile being counted
    int set = 0;
    bool used[2] = { false, false };
    // pending = the tile whose fill has been enqueued and whose sort has not
    bool have_pending = false;
    size_t p_first = 0, p_m = 0; long long p_total = 0; int p_set = 0;
    while (first < n_rays || have_pending) {
        size_t m = 0;
        bool counting = false;
        if (first < n_rays) {
            m = tile < n_rays - first ? tile : n_rays - first;
            // the offsets of this set were last read by the consumer of two tiles ago
            if (used[set]) GB_CUDA(cudaStreamWaitEvent(sa, B[set].consumed, 0));
            rc = hits_count_enqueue(ctx, d_rays + first, m, d_spheres4, n, tree, 0, B[set].offsets, sa);
            if (rc) return cleanup(rc);
            counting = true;
        }
        if (have_pending) {     // sort + consume the previous tile while this one is being counted
            GB_CUDA(cudaStreamWaitEvent(sb, B[p_set].filled, 0));
            if (p_total > 0) {
                rc = grace_b200_sort_by_distance(ctx->aux, B[p_set].dist, B[p_set].offsets, p_m, (size_t)p_total, B[p_set].idx,
                                                 B[p_set].integ, (void*)sb);
                if (rc) return cleanup(rc);
            }
            if (consume) {
                rc = consume(user, p_first, p_m, B[p_set].offsets, p_total, B[p_set].idx, B[p_set].integ, B[p_set].dist, (void*)sb);
                if (rc) return cleanup(gb_set_error(rc, "the tile consumer returned %d", rc));
            }
            GB_CUDA(cudaEventRecord(B[p_set].consumed, sb));
            have_pending = false;
        }
        if (!counting) break;
        long long total = 0;
        if ((rc = hits_count_finish(ctx, &total, sa))) return cleanup(rc);
        if ((unsigned long long)total > hit_budget) {
            if (m <= 32) return cleanup(gb_set_error(GRACE_B200_ERANGE, "one packet of 32 rays has %lld hits: raise the hit budget (%zu)", total, hit_budget));
            tile = (m / 2 + 31) / 32 * 32;       // halve and count again
            continue;
        }
        if (total > 0) {
            rc = grace_b200_trace_hits_fill_f4(ctx, d_rays + first, m, d_spheres4, n, tree, B[set].offsets, B[set].idx, B[set].integ,
                                               B[set].dist, (void*)sa);
            if (rc) return cleanup(rc);
        }
        GB_CUDA(cudaEventRecord(B[set].filled, sa));
        used[set] = true;
        have_pending = true; p_first = first; p_m = m; p_total = total; p_set = set;
        grand += total;
        first += m;
        set ^= 1;
    }
    if (h_total_hits) *h_total_hits = grand;
    return cleanup(GRACE_B200_OK);
}

int grace_b200_sort_by_distance(grace_b200_ctx* ctx, float* d_hit_distances, const int* d_ray_offsets,
                                size_t n_rays, size_t total_hits, int* d_hit_indices,
                                void* d_hit_data, void* stream)
{
    GB_REQUIRE(ctx && d_ray_offsets, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_rays < (1ull << 31) && total_hits < (1ull << 31), GRACE_B200_ERANGE, "too many rays/hits");
    if (n_rays == 0 || total_hits == 0) return GRACE_B200_OK;
    GB_REQUIRE(d_hit_distances && d_hit_indices && d_hit_data, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nr = (int)n_rays;
    const size_t list_bytes = gb_align((size_t)N_CLASSES * n_rays * sizeof(int));
    const size_t xl_off_bytes = gb_align(n_rays * 8);
    char* ws = (char*)gb_workspace(ctx, list_bytes + xl_off_bytes + 256);
    if (!ws) return GRACE_B200_ENOMEM;
    int* lists = (int*)ws;
    unsigned long long* xl_offsets = (unsigned long long*)(ws + list_bytes);
    int* counts = ctx->d_scalars + GB_SC_CLASS;                       // N_CLASS_SLOTS ints
    unsigned long long* xl_total = (unsigned long long*)(ctx->d_scalars + GB_SC_CLASS + N_CLASS_SLOTS);
    GB_CUDA(cudaMemsetAsync(counts, 0, (N_CLASS_SLOTS + 2) * sizeof(int), st));
    classify_kernel<<<(nr + 255) / 256, 256, 0, st>>>(d_ray_offsets, nr, (long long)total_hits, lists,
                                                      counts, xl_total, xl_offsets);
    GB_LAUNCH_CHECK();
    int* h = ctx->h_pinned + GB_SC_CLASS;
    GB_CUDA(cudaMemcpyAsync(h, counts, (N_CLASS_SLOTS + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    float* dist = d_hit_distances;
    int* idx = d_hit_indices;
    unsigned* data = (unsigned*)d_hit_data;
    const long long total = (long long)total_hits;
    int rc;
    static_assert(N_CLASSES == 10 && N_CLASSES <= N_CLASS_SLOTS, "class table below");
#define GB_SORT_CLASS(c, CAP, NT)                                                                         \
    static_assert(class_cap(c) == CAP, "class table out of step with class_cap");                         \
    if ((rc = launch_smem_class<CAP, NT>(ctx, c, dist, d_ray_offsets, nr, total, idx, data, lists, counts, h[c], st))) return rc;
    GB_SORT_CLASS(0, 32, 32)
    GB_SORT_CLASS(1, 512, 128)
    GB_SORT_CLASS(2, 1024, 128)
    GB_SORT_CLASS(3, 1536, 192)
    GB_SORT_CLASS(4, 2048, 256)
    GB_SORT_CLASS(5, 3072, 384)
    GB_SORT_CLASS(6, 4096, 512)
    GB_SORT_CLASS(7, 6144, 768)
    GB_SORT_CLASS(8, 8192, 1024)
#undef GB_SORT_CLASS
    constexpr int XL = N_CLASSES - 1;
    if (h[XL] > 0) {
        const unsigned long long xl = *(unsigned long long*)(h + N_CLASS_SLOTS);
        // the workspace is still holding lists/xl_offsets: extend it without moving them
        const size_t head = list_bytes + xl_off_bytes + 256;
        const size_t need = head + gb_align(xl * 8) + 3 * gb_align(xl * 4) + 256;
        if (need > ctx->ws_bytes) {
            // grow: copy lists through a fresh allocation (rare path)
            void* keep = nullptr;
            GB_CUDA(cudaMalloc(&keep, head));
            GB_CUDA(cudaMemcpyAsync(keep, ws, head, cudaMemcpyDeviceToDevice, st));
            GB_CUDA(cudaStreamSynchronize(st));
            char* nws = (char*)gb_workspace(ctx, need);
            if (!nws) { cudaFree(keep); return GRACE_B200_ENOMEM; }
            GB_CUDA(cudaMemcpyAsync(nws, keep, head, cudaMemcpyDeviceToDevice, st));
            GB_CUDA(cudaStreamSynchronize(st));
            cudaFree(keep);
            ws = nws;
            lists = (int*)ws;
            xl_offsets = (unsigned long long*)(ws + list_bytes);
        }
        char* p = ws + head;
        unsigned long long* comp = (unsigned long long*)p; p += gb_align(xl * 8);
        float* td = (float*)p; p += gb_align(xl * 4);
        int* ti = (int*)p; p += gb_align(xl * 4);
        unsigned* tp = (unsigned*)p;
        int blocks = ctx->sm_count * 2;
        if (blocks > h[XL]) blocks = h[XL];
        segsort_global_kernel<<<blocks, 1024, 0, st>>>(dist, d_ray_offsets, nr, total, idx, data,
                                                       lists + (size_t)XL * n_rays, counts + XL,
                                                       xl_offsets, comp, td, ti, tp);
        GB_LAUNCH_CHECK();
    }
    return GRACE_B200_OK;
}

} // extern "C"
