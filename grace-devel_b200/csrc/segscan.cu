// segscan.cu -- segmented exclusive scans over per-ray hit lists (the step after the hot
// path: cumulative optical depth along each distance-sorted ray), SURVEY.md 8f N3.
//
// Reference behaviour (GRACE): exclusive_segmented_scan, cuda/scan.cuh:15-38 (a new sgpu
// context and a CSR preprocessing pass per call, then sgpu::SegScanApply);
// weighted_exclusive_segmented_scan, cuda/scan.cuh:45-58 (a temporary of N values filled by
// multiply_by_weights_kernel<<<48,512>>>, cuda/kernels/weights.cuh:13-59, then the scan);
// offsets_to_segments, cuda/sort.cuh:20-41 (thrust scatter + inclusive scan).
// Contract: results[i] = sum of data[segment start .. i) ; the reference's own test
// (tests/segmented_scan/segmented_scan.cu:97-140) uses integer-valued data, so the result is
// exact in any association; for general floats the reference's association is sgpu's
// (tile-tree), ours is the one below -- both within rounding of the sequential sum.
//
// B200 design: one pass, no preprocessing, no temporary.  A warp owns a segment (handed out
// by ticket): it walks the segment 32 elements at a time, warp-scans each chunk with
// shuffles and carries the running total in a register, so every element is read once and
// written once with unit stride; the weights are gathered and multiplied in the same pass.
// Segments are ray hit lists (thousands of elements); empty segments cost one ticket.
#include "common.cuh"

namespace {

constexpr int SS_THREADS = 256;

template <bool WEIGHTED>
__global__ void __launch_bounds__(SS_THREADS)
segscan_kernel(const int* __restrict__ offsets, int n_segments, long long n_data,
               const float* data, const float* __restrict__ weights,
               const unsigned* __restrict__ weight_map, float* results, unsigned* ticket)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        int seg = 0;
        if (lane == 0) seg = (int)atomicAdd(ticket, 1u);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= n_segments) break;
        const long long begin = offsets[seg];
        const long long end = seg + 1 < n_segments ? (long long)offsets[seg + 1] : n_data;
        float carry = 0.0f;
        for (long long base = begin; base < end; base += 32) {
            const long long i = base + lane;
            float x = 0.0f;
            if (i < end) {
                x = data[i];
                // weights.cuh:22-23: weighted[i] = weights[weight_map[i]] * unweighted[i]
                if (WEIGHTED) x = __fmul_rn(__ldg(weights + __ldg(weight_map + i)), x);
            }
            float incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (i < end) results[i] = carry + (lane ? excl : 0.0f);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// offsets_to_segments, cuda/sort.cuh:20-41: scatter 1 at offsets[1:], inclusive scan.  With empty
// segments several offsets coincide and the scatter writes ONE 1, so the reference numbers
// the elements by the count of DISTINCT values among offsets[1..s], not by s:
// offsets [0,3,3,7] -> [0,0,0,1,1,1,1,(2)].  flag[k] = 1 iff offsets[k] opens a new value.
__global__ void __launch_bounds__(SS_THREADS)
segment_flags_kernel(const int* __restrict__ offsets, int n_segments, int* __restrict__ flags)
{
    const int k = blockIdx.x * SS_THREADS + threadIdx.x;
    if (k < n_segments) flags[k] = (k >= 1 && (k == 1 || offsets[k] != offsets[k - 1])) ? 1 : 0;
}

__global__ void __launch_bounds__(SS_THREADS)
offsets_to_segments_kernel(const int* __restrict__ offsets, const int* __restrict__ flags,
                           const int* __restrict__ excl, int n_segments, long long n_data,
                           int* __restrict__ segments)
{
    const int warps = gridDim.x * (SS_THREADS / 32);
    const int lane = threadIdx.x & 31;
    for (int seg = blockIdx.x * (SS_THREADS / 32) + (threadIdx.x >> 5); seg < n_segments; seg += warps) {
        const long long begin = offsets[seg];
        const long long end = seg + 1 < n_segments ? (long long)offsets[seg + 1] : n_data;
        const int id = excl[seg] + flags[seg];
        for (long long i = begin + lane; i < end; i += 32) segments[i] = id;
    }
}

template <bool WEIGHTED>
int launch_segscan(grace_b200_ctx* ctx, const int* d_offsets, size_t n_segments, size_t n_data,
                   const float* d_data, const float* d_weights, const unsigned* d_map, float* d_results,
                   cudaStream_t st)
{
    GB_REQUIRE(ctx && d_offsets && d_data && d_results, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_segments < (1ull << 31) && n_data < (1ull << 31), GRACE_B200_ERANGE,
               "segment offsets are 32-bit (cuda/scan.cuh:16)");
    if (n_segments == 0 || n_data == 0) return GRACE_B200_OK;
    // its own counter: the bounds kernel's ticket (TICKET0) resets itself and must be zero between launches
    unsigned* ticket = (unsigned*)(ctx->d_scalars + GB_SC_TICKET3);
    GB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    size_t blocks = (n_segments + SS_THREADS / 32 - 1) / (SS_THREADS / 32);
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    segscan_kernel<WEIGHTED><<<(int)blocks, SS_THREADS, 0, st>>>(d_offsets, (int)n_segments, (long long)n_data,
                                                                  d_data, d_weights, d_map, d_results, ticket);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // namespace

extern "C" {

int grace_b200_exclusive_segmented_scan_f32(grace_b200_ctx* ctx, const int* d_segment_offsets,
                                            size_t n_segments, const float* d_data, size_t n_data,
                                            float* d_results, void* stream)
{
    return launch_segscan<false>(ctx, d_segment_offsets, n_segments, n_data, d_data, nullptr, nullptr,
                                 d_results, (cudaStream_t)stream);
}

int grace_b200_weighted_exclusive_segmented_scan_f32(grace_b200_ctx* ctx, const float* d_to_sum,
                                                     const float* d_weights, const unsigned* d_weight_map,
                                                     const int* d_segment_offsets, size_t n_segments,
                                                     size_t n_data, float* d_sum, void* stream)
{
    GB_REQUIRE(d_weights && d_weight_map, GRACE_B200_EINVAL, "NULL argument");
    return launch_segscan<true>(ctx, d_segment_offsets, n_segments, n_data, d_to_sum, d_weights,
                                d_weight_map, d_sum, (cudaStream_t)stream);
}

int grace_b200_offsets_to_segments(grace_b200_ctx* ctx, const int* d_offsets, size_t n_offsets,
                                   int* d_segments, size_t n_data, void* stream)
{
    GB_REQUIRE(ctx && d_offsets && d_segments, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_offsets < (1ull << 31) && n_data < (1ull << 31), GRACE_B200_ERANGE, "32-bit offsets");
    if (n_offsets == 0 || n_data == 0) return GRACE_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // workspace: [scan state of grace_b200_exclusive_scan_i32 | flags | exclusive counts]; the
    // arena is grow-only and the scan's own (smaller) request returns the same base
    const size_t scan_bytes = gb_align((n_offsets / 2048 + 2) * 8 + 256);
    char* ws = (char*)gb_workspace(ctx, scan_bytes + 2 * gb_align(n_offsets * 4));
    if (!ws) return GRACE_B200_ENOMEM;
    int* flags = (int*)(ws + scan_bytes);
    int* excl = (int*)(ws + scan_bytes + gb_align(n_offsets * 4));
    segment_flags_kernel<<<(int)((n_offsets + SS_THREADS - 1) / SS_THREADS), SS_THREADS, 0, st>>>(
        d_offsets, (int)n_offsets, flags);
    GB_LAUNCH_CHECK();
    int rc = grace_b200_exclusive_scan_i32(ctx, flags, excl, n_offsets, nullptr, stream);
    if (rc) return rc;
    size_t blocks = (n_offsets + SS_THREADS / 32 - 1) / (SS_THREADS / 32);
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    offsets_to_segments_kernel<<<(int)blocks, SS_THREADS, 0, st>>>(d_offsets, flags, excl, (int)n_offsets,
                                                                    (long long)n_data, d_segments);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

} // extern "C"
