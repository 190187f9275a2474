// albvh.cu -- deltas, leaf clustering and bottom-up ALBVH node build.
//
// Reference behaviour (GRACE, include/grace/cuda/kernels/albvh.cuh):
//   compute_deltas_kernel :33-47, build_leaves_kernel :77-234, write_leaves_kernel
//   :236-295, remove_empty_leaves :826-846, copy_leaf_deltas_kernel :51-74,
//   build_nodes_slice_kernel/fill_output_queue/fix_node_ranges :303-761 driven by a
//   host loop with Thrust queue maintenance and a sync per slice (:854-940).
// The finished arrays are a pure function of (sorted spheres, deltas, max_per_leaf):
//   parent rule: a subtree [l, r] is the RIGHT child of node l-1 if
//   delta(l-1) < delta(r), else the LEFT child of node r (ties go right);
//   leaves = maximal subtrees with <= max_per_leaf primitives.
//
// B200 design -- three launches, no host round trip, no Thrust:
//  (1) leaves: node j's subtree is [l_j, r_j] with l_j = 1 + (last k < j with delta_k >= delta_j) and
//      r_j = (first k > j with delta_k > delta_j): the Cartesian tree of the deltas under the tie rule
//      above.  Leaves are the stretches between consecutive nodes whose subtree exceeds max_per_leaf;
//      for max_per_leaf = 32 those nodes are found as sliding-window maxima (leaves_window_kernel),
//      otherwise from bounded scans over a shared-memory table (leaves_kernel).  No atomics, no
//      temporary node array; the leaves (ordered by primitive range) are compacted in the same kernel
//      with a decoupled look-back scan, the leaf-level deltas and the zeroed arrival flags are written
//      alongside.
//  (2) leaf_boxes_kernel: one AABB per leaf, spheres staged through shared memory with coalesced loads.
//  (3) nodes_kernel: subtrees merged in registers and shared memory per warp (no atomics for ~90 % of
//      the nodes), then a Karras/Apetrei bottom-up climb with one global arrival counter per node for
//      what is left, writing child index / range end / AABB into the parent's 64-byte record (the
//      reference layout, cuda/nodes.h:21-36).
// Algorithmic bytes per particle: deltas 16+4; leaves 4 + (16+4+4)*L/N; leaf boxes 16 + 32*L/N;
// nodes (32 + 2*64)*L/N.
#include "common.cuh"

#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int LV_THREADS = 256;
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INCL = 2ull << 62, ST_MASK = 3ull << 62;

// ---------------------------------------------------------------------------
// deltas
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
deltas_euclid_kernel(const float4* __restrict__ s, size_t n, float* __restrict__ deltas)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n; t += stride) {
        float d = CUDART_INF_F;
        if (t >= 1 && t < n) {
            // generic/functors/albvh.h:78-80 as nvcc contracts it (SASS-verified):
            // d = dy*dy; d = fma(dx,dx,d); d = fma(dz,dz,d)
            const float4 a = __ldg(s + t - 1), b = __ldg(s + t);
            const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
            d = __fmul_rn(dy, dy);
            d = __fmaf_rn(dx, dx, d);
            d = __fmaf_rn(dz, dz, d);
        }
        deltas[t] = d;
    }
}

__global__ void __launch_bounds__(256)
deltas_sarea_kernel(const float4* __restrict__ s, size_t n, float* __restrict__ deltas)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n; t += stride) {
        float d = CUDART_INF_F;
        if (t >= 1 && t < n) {
            // generic/functors/albvh.h:101-121: SA = Lx*Lz; fma(Lx,Ly,SA); fma(Ly,Lz,SA)
            const float4 a = __ldg(s + t - 1), b = __ldg(s + t);
            const float Lx = __fsub_rn(fmaxf(__fadd_rn(a.x, a.w), __fadd_rn(b.x, b.w)),
                                       fminf(__fsub_rn(a.x, a.w), __fsub_rn(b.x, b.w)));
            const float Ly = __fsub_rn(fmaxf(__fadd_rn(a.y, a.w), __fadd_rn(b.y, b.w)),
                                       fminf(__fsub_rn(a.y, a.w), __fsub_rn(b.y, b.w)));
            const float Lz = __fsub_rn(fmaxf(__fadd_rn(a.z, a.w), __fadd_rn(b.z, b.w)),
                                       fminf(__fsub_rn(a.z, a.w), __fsub_rn(b.z, b.w)));
            d = __fmul_rn(Lx, Lz);
            d = __fmaf_rn(Lx, Ly, d);
            d = __fmaf_rn(Ly, Lz, d);
        }
        deltas[t] = d;
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(256)
deltas_xor_kernel(const KeyT* __restrict__ keys, size_t n, KeyT* __restrict__ deltas)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n; t += stride) {
        KeyT d = ~KeyT(0);
        if (t >= 1 && t < n) d = keys[t - 1] ^ keys[t];
        deltas[t] = d;
    }
}

// ---------------------------------------------------------------------------
// leaves
// ---------------------------------------------------------------------------
// d(k) = deltas_shifted[k + 1] is the delta between primitives k and k+1, valid for
// k in [-1, n-1]; d(-1) and d(n-1) are the sentinels.
//
// One block handles LV_TILE consecutive nodes, LV_ITEMS consecutive nodes per thread, so a
// thread's (0..2 per node) leaves are contiguous in the output and one block-wide scan of the
// per-thread totals orders them.  The tile's running offset comes from a decoupled look-back
// in which a whole warp inspects 32 predecessor tiles per step.
constexpr int LV_ITEMS = 8;
constexpr int LV_TILE = LV_THREADS * LV_ITEMS;

// window index -> shared-memory index: one pad word per 32 keeps the LV_ITEMS-strided
// accesses of a warp on distinct banks
__device__ __forceinline__ int lv_sw(int i) { return i + (i >> 5); }

// K > 0: the two bounded scans ("how many consecutive neighbours are smaller") are answered from
// a sparse table of window maxima, M_k[i] = max d[i .. i + 2^k), by binary descent: K steps per
// side whatever the data, where the plain scan costs a warp the LONGEST scan among its lanes
// (close to the cap mpl + 1 for most warps).  2^K > mpl.  K = 0: plain scans (large mpl).
template <typename T, bool STAGED, int K>
__global__ void __launch_bounds__(LV_THREADS)
leaves_kernel(const T* __restrict__ deltas_shifted, int n, int mpl,
              int4* __restrict__ leaves, T* __restrict__ leaf_deltas_shifted,
              unsigned* __restrict__ node_flags, unsigned long long* __restrict__ block_state,
              unsigned* __restrict__ ticket, int* __restrict__ n_leaves_out)
{
    extern __shared__ __align__(16) unsigned char lv_smem[];
    T* win = (T*)lv_smem;
    __shared__ unsigned s_bid;
    __shared__ unsigned s_warp_tot[LV_THREADS / 32];
    __shared__ unsigned long long s_excl;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const int n_nodes = n - 1;
    const int j0 = (int)bid * LV_TILE;
    const int pad = mpl + 2;
    // window covers k in [j0 - pad, j0 + LV_TILE + pad)
    const int wlo = j0 - pad;
    const int wn = LV_TILE + 2 * pad;
    const int wstride = lv_sw(wn) + 1;       // words per table level
    if (STAGED) {
        for (int i = tid; i < wn; i += LV_THREADS) {
            int k = wlo + i;
            k = max(k, -1);
            k = min(k, n - 1);
            win[lv_sw(i)] = deltas_shifted[k + 1];
        }
        __syncthreads();
#pragma unroll
        for (int k = 1; k < K; ++k) {
            const T* src = win + (k - 1) * wstride;
            T* dst = win + k * wstride;
            const int half = 1 << (k - 1);
            for (int i = tid; i < wn; i += LV_THREADS) {
                const T a = src[lv_sw(i)];
                const T b = src[lv_sw(min(i + half, wn - 1))];
                dst[lv_sw(i)] = a < b ? b : a;
            }
            __syncthreads();
        }
    }
    auto d = [&](int k) -> T {
        if (STAGED) return win[lv_sw(k - wlo)];
        return deltas_shifted[k + 1];
    };

    // per node: l, r of its subtree (capped scans) and which children are leaves
    int lo[LV_ITEMS], hi[LV_ITEMS];
    unsigned emit_bits = 0;           // bit 2i: left child of node i is a leaf, bit 2i+1: right child
    int emit = 0;
#pragma unroll
    for (int i = 0; i < LV_ITEMS; ++i) {
        const int j = j0 + tid * LV_ITEMS + i;
        int l = j, r = j + 1;
        if (j < n_nodes) {
            const T dj = d(j);
            if (K > 0) {
                // left: longest run d(j-1), d(j-2), ... < dj, at most min(mpl, j) long
                const int capl = min(mpl, j), capr = min(mpl, n - 2 - j);
                int m = 0;
                int pos = j - wlo;               // window index one past the run's low end
#pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    const int step = 1 << k;
                    if (m + step <= capl && win[k * wstride + lv_sw(pos - step)] < dj) { pos -= step; m += step; }
                }
                l = j - m;
                // right: longest run d(j+1), d(j+2), ... <= dj, at most min(mpl, n-2-j) long
                m = 0;
                pos = j + 1 - wlo;
#pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    const int step = 1 << k;
                    if (m + step <= capr && !(dj < win[k * wstride + lv_sw(pos)])) { pos += step; m += step; }
                }
                r = j + 1 + m;
            } else {
                while (l > 0 && (j - l + 1) <= mpl && d(l - 1) < dj) --l;
                while (r < n - 1 && (r - j) <= mpl && !(dj < d(r))) ++r;
            }
            const int left_size = j - l + 1, right_size = r - j;
            const bool big = left_size + right_size > mpl;   // sizes are capped at mpl+1
            const bool eL = big && left_size <= mpl, eR = big && right_size <= mpl;
            emit_bits |= ((unsigned)eL << (2 * i)) | ((unsigned)eR << (2 * i + 1));
            emit += (int)eL + (int)eR;
        }
        lo[i] = l; hi[i] = r;
    }
    // block-wide exclusive scan of emit
    unsigned incl = emit;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    unsigned add = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < LV_THREADS / 32; ++w) {
        if (w < warp) add += s_warp_tot[w];
        block_total += s_warp_tot[w];
    }
    const unsigned local_excl = add + incl - emit;

    if (warp == 0) {
        unsigned long long excl = 0;
        if (bid == 0) {
            if (lane == 0) gb_st_volatile_u64(block_state, (unsigned long long)block_total | ST_INCL);
        } else {
            if (lane == 0) gb_st_volatile_u64(block_state + bid, (unsigned long long)block_total | ST_AGG);
            int t = (int)bid - 1;           // nearest predecessor not yet accounted for
            for (;;) {
                const int idx = t - lane;
                unsigned long long v = ST_INCL;      // tiles before the first contribute 0
                if (idx >= 0) {
                    do { v = gb_ld_volatile_u64(block_state + idx); } while ((v & ST_MASK) == 0);
                }
                const unsigned im = __ballot_sync(0xffffffffu, (v & ST_MASK) == ST_INCL);
                const int first = im ? __ffs(im) - 1 : 32;       // nearest tile with an inclusive prefix
                unsigned long long c = lane <= first ? (v & ~ST_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (im) break;
                t -= 32;
            }
            if (lane == 0) gb_st_volatile_u64(block_state + bid, (excl + block_total) | ST_INCL);
        }
        if (lane == 0) {
            s_excl = excl;
            if ((int)bid == (n_nodes - 1) / LV_TILE) *n_leaves_out = (int)(excl + block_total);
            if (bid == 0) leaf_deltas_shifted[0] = deltas_shifted[0];
        }
    }
    __syncthreads();
    if (emit) {
        unsigned g = (unsigned)s_excl + local_excl;
#pragma unroll
        for (int i = 0; i < LV_ITEMS; ++i) {
            const int j = j0 + tid * LV_ITEMS + i;
            if (emit_bits & (1u << (2 * i))) {
                leaves[g] = make_int4(lo[i], j - lo[i] + 1, 0, 0);
                leaf_deltas_shifted[g + 1] = d(j);          // delta after the leaf's last primitive
                node_flags[g] = 0u;
                ++g;
            }
            if (emit_bits & (1u << (2 * i + 1))) {
                leaves[g] = make_int4(j + 1, hi[i] - j, 0, 0);
                leaf_deltas_shifted[g + 1] = d(hi[i]);
                node_flags[g] = 0u;
                ++g;
            }
        }
    }
}

// max_per_leaf == W: leaves from a sliding-window argmax, no per-node scans at all.
//
// Call a node "big" when its subtree holds more than W primitives.  The leaves are exactly the
// stretches between consecutive big nodes (a leaf [l, k] ends at the big node k that is its parent or
// its parent's ... -- every node strictly inside is small, and so is the stretch's own root, whose
// range is the whole stretch), with virtual big nodes at -1 and n - 1.  And node k is big iff the
// interval around k on which delta(k) is the maximum -- left neighbours strictly smaller, right
// neighbours smaller or equal, the tie rule of the reference -- spans more than W primitives, i.e. iff
// k is the LEFTMOST maximum of some window of W consecutive deltas inside [0, n - 2].  Sliding-window
// maxima cost O(1) per element (van Herk / Gil-Werman): with the deltas cut into blocks of W, the
// window starting at element i of block b is the suffix [i, W) of block b + the prefix [0, i) of block
// b + 1, so one thread owning a block's suffix maxima and the next block's prefix maxima marks the
// argmax of W windows with ~14 instructions per node, all in registers (the table descent of
// leaves_kernel took ~300 and was issue-bound at 3 % of the DRAM bandwidth).
//
// Thread t of a CTA works on the windows starting in block B0 - 2 + t; the marks falling into the next
// block reach its owner through shared memory, so blocks B0 - 1 .. B0 + 253 have complete masks and
// the CTA emits the leaves ending in blocks B0 .. B0 + 253 (the block before supplies "previous big
// node").  Compaction, look-back and output are those of leaves_kernel.
constexpr int LW_OWNED = LV_THREADS - 2;      // blocks of W nodes a CTA emits

// (five CTAs per SM for 32-bit deltas, 51 registers: 80 -> 70 us; the kernel waits on one staging round trip per tile)
template <typename T, int W>
__global__ void __launch_bounds__(LV_THREADS, sizeof(T) == 4 ? 5 : 1)
leaves_window_kernel(const T* __restrict__ deltas_shifted, int n,
                     int4* __restrict__ leaves, T* __restrict__ leaf_deltas_shifted,
                     unsigned* __restrict__ node_flags, unsigned long long* __restrict__ block_state,
                     unsigned* __restrict__ ticket, int* __restrict__ n_leaves_out, int n_tiles)
{
    extern __shared__ __align__(16) unsigned char lw_smem[];
    T* win = (T*)lw_smem;                         // (LV_THREADS + 1) blocks, stride W + 1
    __shared__ unsigned s_next[LV_THREADS];       // marks a thread's windows left in the following block
    __shared__ unsigned s_mask[LV_THREADS];       // complete mask of the thread's block
    __shared__ unsigned s_bid;
    __shared__ unsigned s_warp_tot[LV_THREADS / 32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const long long first_block = (long long)bid * LW_OWNED - 2;       // thread 0's block
    // stage deltas of positions [first_block * W, (first_block + LV_THREADS + 1) * W): coalesced, clamped
    {
        const long long p0 = first_block * W;
        for (int i = tid; i < (LV_THREADS + 1) * W; i += LV_THREADS) {
            long long k = p0 + i;
            k = k < -1 ? -1 : (k > n - 1 ? n - 1 : k);
            win[(i / W) * (W + 1) + (i % W)] = __ldg(deltas_shifted + k + 1);
        }
    }
    __syncthreads();
    const long long blk = first_block + tid;      // this thread's block; its positions are blk * W + [0, W)
    const T* mine = win + tid * (W + 1);
    const T* next = mine + (W + 1);
    // prefix maxima of the next block (value + leftmost position, as a mask of strict records)
    T pv[W];
    unsigned rec = 1u;
    pv[0] = next[0];
#pragma unroll
    for (int i = 1; i < W; ++i) {
        const T v = next[i];
        const bool up = pv[i - 1] < v;
        pv[i] = up ? v : pv[i - 1];
        rec |= up ? (1u << i) : 0u;
    }
    // windows by descending start: running suffix maximum of this block (leftmost on ties)
    unsigned A = 0u, Bn = 0u;
    {
        const long long s_max = (long long)n - 1 - W;          // last valid window start
        T sv = mine[W - 1];
        int si = W - 1;
#pragma unroll
        for (int i = W - 1; i >= 0; --i) {
            if (i < W - 1) {
                const T v = mine[i];
                if (!(v < sv)) { sv = v; si = i; }
            }
            const long long s = blk * W + i;
            if (s >= 0 && s <= s_max) {
                if (i == 0 || !(sv < pv[i > 0 ? i - 1 : 0])) A |= 1u << si;
                else Bn |= 1u << (31 - __clz(rec & ((1u << i) - 1u)));
            }
        }
    }
    s_next[tid] = Bn;
    __syncthreads();
    unsigned M = A | (tid > 0 ? s_next[tid - 1] : 0u);
    // virtual big node at n - 1 (end of the last leaf)
    if ((long long)(n - 1) >= blk * W && (long long)(n - 1) < (blk + 1) * W) M |= 1u << (int)((n - 1) - blk * W);
    s_mask[tid] = M;
    __syncthreads();
    const bool owner = tid >= 2 && blk * W <= (long long)(n - 1);
    const unsigned emit = owner ? __popc(M) : 0u;
    unsigned incl = emit;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    unsigned add = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < LV_THREADS / 32; ++w) {
        if (w < warp) add += s_warp_tot[w];
        block_total += s_warp_tot[w];
    }
    const unsigned local_excl = add + incl - emit;
    if (warp == 0) {
        unsigned long long excl = 0;
        if (bid == 0) {
            if (lane == 0) gb_st_volatile_u64(block_state, (unsigned long long)block_total | ST_INCL);
        } else {
            if (lane == 0) gb_st_volatile_u64(block_state + bid, (unsigned long long)block_total | ST_AGG);
            int t = (int)bid - 1;           // nearest predecessor not yet accounted for
            for (;;) {
                const int idx = t - lane;
                unsigned long long v = ST_INCL;      // tiles before the first contribute 0
                if (idx >= 0) {
                    do { v = gb_ld_volatile_u64(block_state + idx); } while ((v & ST_MASK) == 0);
                }
                const unsigned im = __ballot_sync(0xffffffffu, (v & ST_MASK) == ST_INCL);
                const int first = im ? __ffs(im) - 1 : 32;       // nearest tile with an inclusive prefix
                unsigned long long cc = lane <= first ? (v & ~ST_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, o);
                excl += cc;
                if (im) break;
                t -= 32;
            }
            if (lane == 0) gb_st_volatile_u64(block_state + bid, (excl + block_total) | ST_INCL);
        }
        if (lane == 0) {
            s_excl = excl;
            if ((int)bid == n_tiles - 1) *n_leaves_out = (int)(excl + block_total);
            if (bid == 0) leaf_deltas_shifted[0] = deltas_shifted[0];
        }
    }
    __syncthreads();
    if (emit) {
        unsigned g = (unsigned)s_excl + local_excl;
        // the big node before this block's first: in the previous block (a leaf holds at most W primitives), or -1
        const unsigned pm = s_mask[tid - 1];
        long long prev = pm ? (blk - 1) * W + (31 - __clz(pm)) : -1;
        unsigned m = M;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const long long k = blk * W + bit;
            leaves[g] = make_int4((int)(prev + 1), (int)(k - prev), 0, 0);
            leaf_deltas_shifted[g + 1] = mine[bit];         // delta after the leaf's last primitive
            node_flags[g] = 0u;
            prev = k;
            ++g;
        }
    }
}

// ---------------------------------------------------------------------------
// nodes
// ---------------------------------------------------------------------------
// nodes_kernel builds the inner nodes bottom-up in three stages of decreasing locality:
//  (1) one lane per leaf AABB;
//  (2) merging in registers: the finished subtrees a warp holds tile a range of leaves in order,
//      one per active lane.  [l, r] is the right child of node l-1 if delta(l-1) < delta(r), else
//      the left child of node r (ties go right), so a LEFT child whose next subtree is a RIGHT
//      child has found its sibling (both name node r): the left lane fetches the sibling's state
//      by shuffle, writes the parent's complete 64-byte record (the reference layout,
//      cuda/nodes.h:21-36) and becomes the parent; the right lane retires.  Rounds repeat until
//      nothing merges.  Applied first to the 32 leaves of each warp, then -- through shared
//      memory -- to what the 8 warps of the block have left (256 consecutive leaves).  No atomics,
//      no fences: ~99 % of the nodes;
//  (3) the few subtrees a block cannot finish climb Karras/Apetrei-style through one arrival
//      counter per node (fence + atomic per level: this was 2/3 of the kernel when every warp's
//      leftovers went straight to it).
constexpr int ND_THREADS = 256;
constexpr int ND_WARPS = ND_THREADS / 32;

template <typename T>
struct NdSub {            // a finished subtree
    int l, r, cur;        // leaf range, node index (>= n_nodes: leaf)
    T dl, dr;             // delta(l - 1), delta(r)
    float bx, by, bz, tx, ty, tz;
};

// One warp merges the subtrees its active lanes hold (in order, tiling a leaf range).
template <typename T>
__device__ __forceinline__ void nd_merge_rounds(NdSub<T>& S, bool& active, int lane, int4* nodes)
{
    for (;;) {
        const bool right_child = S.dl < S.dr;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        const unsigned higher = lane == 31 ? 0u : (amask & ~((2u << lane) - 1u));
        const int nxt = higher ? __ffs(higher) - 1 : -1;
        const int sl = nxt >= 0 ? nxt : lane;
        const bool s_right = __shfl_sync(0xffffffffu, (int)right_child, sl);
        const int s_r = __shfl_sync(0xffffffffu, S.r, sl);
        const int s_cur = __shfl_sync(0xffffffffu, S.cur, sl);
        const T s_dr = __shfl_sync(0xffffffffu, S.dr, sl);
        const float sbx = __shfl_sync(0xffffffffu, S.bx, sl), sby = __shfl_sync(0xffffffffu, S.by, sl),
                    sbz = __shfl_sync(0xffffffffu, S.bz, sl), stx = __shfl_sync(0xffffffffu, S.tx, sl),
                    sty = __shfl_sync(0xffffffffu, S.ty, sl), stz = __shfl_sync(0xffffffffu, S.tz, sl);
        const bool merge = active && !right_child && nxt >= 0 && s_right;
        const unsigned absorbed = __reduce_or_sync(0xffffffffu, merge ? (1u << sl) : 0u);
        if (absorbed == 0u) break;
        if (merge) {
            const int parent = S.r;
            int4* np = nodes + 4 * (size_t)parent;
            np[0] = make_int4(S.cur, s_cur, S.l, s_r);
            np[1] = make_int4(__float_as_int(S.bx), __float_as_int(S.tx), __float_as_int(S.by), __float_as_int(S.ty));
            np[2] = make_int4(__float_as_int(sbx), __float_as_int(stx), __float_as_int(sby), __float_as_int(sty));
            np[3] = make_int4(__float_as_int(S.bz), __float_as_int(S.tz), __float_as_int(sbz), __float_as_int(stz));
            S.cur = parent; S.r = s_r; S.dr = s_dr;
            S.bx = fminf(S.bx, sbx); S.by = fminf(S.by, sby); S.bz = fminf(S.bz, sbz);
            S.tx = fmaxf(S.tx, stx); S.ty = fmaxf(S.ty, sty); S.tz = fmaxf(S.tz, stz);
        }
        if ((absorbed >> lane) & 1u) active = false;
    }
}

// Leaf boxes, as their own pass: one lane per leaf reading its spheres straight from global memory touches 32
// cache lines per load instruction (the L1 tag stage alone was ~100 us of the node build), so a warp stages the
// contiguous sphere range of its 32 leaves in shared memory with coalesced cp.async, LB_SLAB spheres at a time,
// and every lane folds the part of its own leaf that lies in the slab.  Boxes go to a scratch array
// ({bx,by,bz,tx} {ty,tz,-,-} per leaf); the node build reads one lane's box with two coalesced loads.
// FROM_AABB: `prims` holds two float4 per primitive, {bx,by,bz,-} {tx,ty,tz,-}: boxes a user's AABB functor
// produced (generic primitives), instead of {x,y,z,h} spheres.
constexpr int LB_THREADS = 256;
constexpr int LB_SLAB = 512;          // 256: +3 us, 1024: +17 us

template <bool FROM_AABB>
__global__ void __launch_bounds__(LB_THREADS)
leaf_boxes_kernel(const float4* __restrict__ prims, const int4* __restrict__ leaves, const int* __restrict__ n_leaves_ptr,
                  float4* __restrict__ boxes)
{
    extern __shared__ __align__(16) unsigned char lb_smem[];
    constexpr int PER = FROM_AABB ? 2 : 1;           // float4 per primitive
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* slab = (float4*)lb_smem + (size_t)warp * LB_SLAB * PER;
    const int L = *n_leaves_ptr;
    const int rows = (L + 31) / 32;
    const int n_warps = gridDim.x * (LB_THREADS / 32);
    for (int row = blockIdx.x * (LB_THREADS / 32) + warp; row < rows; row += n_warps) {
        const int leaf = row * 32 + lane;
        const bool active = leaf < L;
        int2 lf = make_int2(0, 0);
        if (active) lf = __ldg((const int2*)(leaves + leaf));
        // the row's leaves tile one contiguous range of primitives
        const int s0 = __shfl_sync(0xffffffffu, lf.x, 0);
        const int s1 = __shfl_sync(0xffffffffu, lf.x + lf.y, min(31, L - 1 - row * 32));
        float bx = CUDART_INF_F, by = CUDART_INF_F, bz = CUDART_INF_F;
        float tx = -CUDART_INF_F, ty = -CUDART_INF_F, tz = -CUDART_INF_F;
        for (int p0 = s0; p0 < s1; p0 += LB_SLAB) {
            const int n = min(LB_SLAB, s1 - p0);
            __syncwarp();              // the previous slab has been read
            const float4* src = prims + (size_t)p0 * PER;
            for (int i = lane; i < n * PER; i += 32) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(slab + i);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src + i) : "memory");
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int lo = max(lf.x, p0) - p0, hi = min(lf.x + lf.y, p0 + n) - p0;
#pragma unroll 4
            for (int i = lo; i < hi; ++i) {
                if (FROM_AABB) {
                    const float4 b = slab[2 * i], t = slab[2 * i + 1];
                    bx = fminf(bx, b.x); by = fminf(by, b.y); bz = fminf(bz, b.z);
                    tx = fmaxf(tx, t.x); ty = fmaxf(ty, t.y); tz = fmaxf(tz, t.z);
                } else {
                    const float4 sp = slab[i];
                    // AABBSphere, generic/functors/aabb.h:9-26: centre -/+ h, one FADD each
                    bx = fminf(bx, __fsub_rn(sp.x, sp.w)); tx = fmaxf(tx, __fadd_rn(sp.x, sp.w));
                    by = fminf(by, __fsub_rn(sp.y, sp.w)); ty = fmaxf(ty, __fadd_rn(sp.y, sp.w));
                    bz = fminf(bz, __fsub_rn(sp.z, sp.w)); tz = fmaxf(tz, __fadd_rn(sp.z, sp.w));
                }
            }
        }
        if (active) {
            boxes[2 * (size_t)leaf] = make_float4(bx, by, bz, tx);
            boxes[2 * (size_t)leaf + 1] = make_float4(ty, tz, 0.f, 0.f);
        }
    }
}

// Node build from the leaf boxes (leaf_boxes_kernel).
//
// Warps are independent (no block barrier: with the block-wide stage the kernel spent its time waiting
// for the one warp that merged the block's leftovers and for the lanes climbing with atomics): a warp
// takes a ticket for ND_ROWS x 32 consecutive leaves, folds and merges them row by row, collects what
// is left in its own shared-memory buffer, merges that, and climbs with the rest.
constexpr int ND_ROWS = 4;            // 2 and 8: +5-10 us
constexpr int ND_GROUP = 32 * ND_ROWS;

template <typename T>
__global__ void __launch_bounds__(ND_THREADS)
nodes_kernel(const float4* __restrict__ boxes, const int* __restrict__ n_leaves_ptr, const T* __restrict__ ld_shifted,
             int4* nodes, unsigned* flags, int* __restrict__ root, int* __restrict__ ticket)
{
    extern __shared__ __align__(16) unsigned char nd_smem[];
    const int L = *n_leaves_ptr;
    const int n_nodes = L - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    NdSub<T>* buf = (NdSub<T>*)nd_smem + warp * ND_GROUP;
    const int n_groups = (L + ND_GROUP - 1) / ND_GROUP;
    for (;;) {
    int grp = 0;
    if (lane == 0) grp = atomicAdd(ticket, 1);
    grp = __shfl_sync(0xffffffffu, grp, 0);
    if (grp >= n_groups) break;
    const int G = grp * ND_GROUP;
    // the leaf boxes of all rows are requested first
    float4 bxs[ND_ROWS][2];
#pragma unroll
    for (int row = 0; row < ND_ROWS; ++row) {
        const int leaf = G + row * 32 + lane;
        bxs[row][0] = bxs[row][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (leaf < L) { bxs[row][0] = __ldg(boxes + 2 * (size_t)leaf); bxs[row][1] = __ldg(boxes + 2 * (size_t)leaf + 1); }
    }
    int count = 0;
#pragma unroll
    for (int row = 0; row < ND_ROWS; ++row) {
        const int leaf = G + row * 32 + lane;
        bool active = leaf < L;
        if (!__any_sync(0xffffffffu, active)) break;
        // ---- (1) leaf boxes ----
        NdSub<T> S;
        S.bx = bxs[row][0].x; S.by = bxs[row][0].y; S.bz = bxs[row][0].z;
        S.tx = bxs[row][0].w; S.ty = bxs[row][1].x; S.tz = bxs[row][1].y;
        // ---- (2a) the row's 32 leaves ----
        S.l = S.r = leaf;
        S.cur = leaf + n_nodes;            // child index >= n_nodes marks a leaf
        S.dl = ld_shifted[min(leaf, L)];            // delta(leaf - 1)
        S.dr = ld_shifted[min(leaf + 1, L)];        // delta(leaf)
        nd_merge_rounds<T>(S, active, lane, nodes);
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) buf[count + __popc(amask & lt)] = S;
        count += __popc(amask);
        __syncwarp();
    }
    // ---- (2b) what the rows have left, in order ----
    // sweeps over windows of 32 subtrees; survivors are compacted to the front.  A pair straddling a
    // window boundary may stay unmerged: stage (3) takes whatever is left.
    for (int sweep = 0; sweep < 4 && count > 1; ++sweep) {
        int out = 0;
        bool merged_any = false;
        for (int wb = 0; wb < count; wb += 32) {
            bool act = wb + lane < count;
            NdSub<T> X = buf[min(wb + lane, ND_GROUP - 1)];
            const int before = __popc(__ballot_sync(0xffffffffu, act));
            nd_merge_rounds<T>(X, act, lane, nodes);
            const unsigned am = __ballot_sync(0xffffffffu, act);
            merged_any |= __popc(am) != before;
            __syncwarp();
            if (act) buf[out + __popc(am & lt)] = X;     // out <= wb: never ahead of the reads
            out += __popc(am);
            __syncwarp();
        }
        count = out;
        if (!merged_any) break;
    }
    // ---- (3) climb across warps ----
    for (int cb = 0; cb < count; cb += 32) {
    bool active = cb + lane < count;
    NdSub<T> S = buf[min(cb + lane, ND_GROUP - 1)];
    int cur = S.cur, l = S.l, r = S.r;
    float bx = S.bx, by = S.by, bz = S.bz, tx = S.tx, ty = S.ty, tz = S.tz;
    bool right_child = active && S.dl < S.dr;
    int parent = right_child ? l - 1 : r;
    while (active) {
        // parent / right_child are current for [l, r] on entry
        if (parent < 0 || parent >= n_nodes) { *root = cur; break; }
        int* node_i = (int*)(nodes + 4 * (size_t)parent);
        float* node_f = (float*)node_i;
        if (right_child) {
            node_i[1] = cur; node_i[3] = r;
            *(float4*)(node_f + 8) = make_float4(bx, tx, by, ty);
            *(float2*)(node_f + 14) = make_float2(bz, tz);
        } else {
            node_i[0] = cur; node_i[2] = l;
            *(float4*)(node_f + 4) = make_float4(bx, tx, by, ty);
            *(float2*)(node_f + 12) = make_float2(bz, tz);
        }
        __threadfence();
        if (atomicAdd(flags + parent, 1u) == 0u) break;    // first arrival stops
        __threadfence();
        const int4 n0 = gb_ld_cg_i4(nodes + 4 * (size_t)parent + 0);
        const int4 n1 = gb_ld_cg_i4(nodes + 4 * (size_t)parent + 1);
        const int4 n2 = gb_ld_cg_i4(nodes + 4 * (size_t)parent + 2);
        const int4 n3 = gb_ld_cg_i4(nodes + 4 * (size_t)parent + 3);
        cur = parent;
        l = n0.z; r = n0.w;
        bx = fminf(__int_as_float(n1.x), __int_as_float(n2.x));
        tx = fmaxf(__int_as_float(n1.y), __int_as_float(n2.y));
        by = fminf(__int_as_float(n1.z), __int_as_float(n2.z));
        ty = fmaxf(__int_as_float(n1.w), __int_as_float(n2.w));
        bz = fminf(__int_as_float(n3.x), __int_as_float(n3.z));
        tz = fmaxf(__int_as_float(n3.y), __int_as_float(n3.w));
        const T dl = ld_shifted[l];          // delta(l - 1)
        const T dr = ld_shifted[r + 1];      // delta(r)
        right_child = dl < dr;
        parent = right_child ? l - 1 : r;
    }
    }
    __syncwarp();       // the buffer is reused by the next group
    }
}

int grid_cap(const grace_b200_ctx* ctx, size_t work_items, int per_block, int per_sm)
{
    size_t blocks = (work_items + per_block - 1) / per_block;
    const size_t cap = (size_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Workspace of one build: leaf-level deltas, per-node arrival flags, look-back states.
template <typename T>
struct BuildWs { T* leaf_deltas; unsigned* flags; unsigned long long* block_state; float4* leaf_boxes; };

template <typename T>
int build_workspace(grace_b200_ctx* ctx, size_t n, BuildWs<T>* out)
{
    const int lv_blocks = ((int)n - 1 + LV_TILE - 1) / LV_TILE;
    // (leaf boxes: 32 bytes per leaf, and the leaf count is only known on the device -- anything up to n; the arena of the
    // keys + sort call that normally precedes is larger)
    const size_t bytes = gb_align((n + 1) * sizeof(T)) + gb_align(n * sizeof(unsigned)) +
                         gb_align((size_t)lv_blocks * 8) + gb_align(2 * n * sizeof(float4)) + 256;
    void* ws = gb_workspace(ctx, bytes);
    if (!ws) return GRACE_B200_ENOMEM;
    GbArena a(ws, bytes);
    out->leaf_deltas = a.take<T>(n + 1);
    out->flags = a.take<unsigned>(n);
    out->block_state = a.take<unsigned long long>(lv_blocks);
    out->leaf_boxes = a.take<float4>(2 * n);
    return GRACE_B200_OK;
}

// Stage 1 (build_leaves + remove_empty_leaves + copy_leaf_deltas): dense leaves, their count in
// d_scalars[GB_SC_NLEAVES], leaf-level deltas and zeroed arrival flags in the workspace.
template <typename T>
int leaves_stage(grace_b200_ctx* ctx, size_t n, const T* d_deltas, int mpl, int4* d_leaves, const BuildWs<T>& w,
                 cudaStream_t st)
{
    const int n_nodes = (int)n - 1;
    const int lv_blocks = (n_nodes + LV_TILE - 1) / LV_TILE;
    unsigned* ticket = (unsigned*)(ctx->d_scalars + GB_SC_TICKET1);
    int* d_nleaves = ctx->d_scalars + GB_SC_NLEAVES;
    GB_CUDA(cudaMemsetAsync(w.block_state, 0, (size_t)lv_blocks * 8, st));
    GB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    // table levels: 2^K > mpl; the table must fit in shared memory
    int K = 0;
    while ((1 << K) <= mpl) ++K;
    const size_t wn = (size_t)LV_TILE + 2 * (mpl + 2);
    const size_t level_words = wn + wn / 32 + 2;
    auto launch = [&](auto kernel, size_t smem) -> int {
        GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<lv_blocks, LV_THREADS, smem, st>>>(d_deltas, (int)n, mpl, d_leaves, w.leaf_deltas, w.flags,
                                                    w.block_state, ticket, d_nleaves);
        return GRACE_B200_OK;
    };
    int lrc;
    static const bool table_only = getenv("GRACE_B200_LEAVES_TABLE") != nullptr;      // A/B switch: the shared-memory table path
    if (mpl == 32 && !table_only) {
        const int n_blocks = ((int)n + 31) / 32;              // positions 0 .. n - 1
        const int n_tiles = (n_blocks + LW_OWNED - 1) / LW_OWNED;          // <= lv_blocks: block_state is large enough
        const size_t smem = sizeof(T) * (LV_THREADS + 1) * 33;
        auto kernel = leaves_window_kernel<T, 32>;
        GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<n_tiles, LV_THREADS, smem, st>>>(d_deltas, (int)n, d_leaves, w.leaf_deltas, w.flags, w.block_state, ticket,
                                                  d_nleaves, n_tiles);
        lrc = GRACE_B200_OK;
    } else if (K <= 6 && sizeof(T) * 6 * level_words <= 112 * 1024) lrc = launch(leaves_kernel<T, true, 6>, sizeof(T) * 6 * level_words);
    else if (K <= 8 && sizeof(T) * 8 * level_words <= 112 * 1024) lrc = launch(leaves_kernel<T, true, 8>, sizeof(T) * 8 * level_words);
    else if (mpl <= 2048) lrc = launch(leaves_kernel<T, true, 0>, sizeof(T) * level_words);
    else lrc = launch(leaves_kernel<T, false, 0>, 0);
    if (lrc) return lrc;
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

// Stage 2 (build_nodes): `cap` bounds the leaf count, which is read on the device (d_nleaves).  `boxes` is scratch for
// two float4 per leaf.
template <typename T>
int nodes_stage(grace_b200_ctx* ctx, const float4* d_prims, bool from_aabb, size_t cap, const int4* d_leaves,
                const int* d_nleaves, const T* leaf_deltas, unsigned* flags, float4* boxes, int4* d_nodes, int* d_root,
                cudaStream_t st)
{
    // (1) leaf boxes: warps stride over rows of 32 leaves
    {
        const size_t smem = (size_t)(LB_THREADS / 32) * LB_SLAB * (from_aabb ? 2 : 1) * sizeof(float4);
        auto launch = [&](auto kernel) -> int {
            GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, LB_THREADS, smem));
            const int need = (int)std::min<size_t>((cap + LB_THREADS - 1) / LB_THREADS, 1u << 30);
            const int blocks = std::max(1, std::min(need, ctx->sm_count * std::max(per_sm, 1)));
            kernel<<<blocks, LB_THREADS, smem, st>>>(d_prims, d_leaves, d_nleaves, boxes);
            return GRACE_B200_OK;
        };
        const int lrc = from_aabb ? launch(leaf_boxes_kernel<true>) : launch(leaf_boxes_kernel<false>);
        if (lrc) return lrc;
        GB_LAUNCH_CHECK();
    }
    // (2) nodes: warps take tickets for ND_GROUP-leaf groups; the grid is what the SMs can hold
    int* nd_ticket = ctx->d_scalars + GB_SC_TICKET2;
    GB_CUDA(cudaMemsetAsync(nd_ticket, 0, sizeof(int), st));
    const size_t smem = (size_t)ND_WARPS * ND_GROUP * sizeof(NdSub<T>);
    auto kernel = nodes_kernel<T>;
    GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ND_THREADS, smem));
    const int need = (int)std::min<size_t>((cap + (size_t)ND_WARPS * ND_GROUP - 1) / ((size_t)ND_WARPS * ND_GROUP), 1u << 30);
    const int nd_blocks = std::max(1, std::min(need, ctx->sm_count * std::max(per_sm, 1)));
    kernel<<<nd_blocks, ND_THREADS, smem, st>>>(boxes, d_nleaves, leaf_deltas, d_nodes, flags, d_root, nd_ticket);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

template <typename T>
int build_typed(grace_b200_ctx* ctx, const float4* d_spheres, size_t n, const T* d_deltas,
                int mpl, int4* d_nodes, int4* d_leaves, int* d_root, cudaStream_t st, bool from_aabb = false)
{
    BuildWs<T> w;
    int rc = build_workspace<T>(ctx, n, &w);
    if (rc) return rc;
    if ((rc = leaves_stage<T>(ctx, n, d_deltas, mpl, d_leaves, w, st))) return rc;
    // the leaf count is only known on the device (anything up to n)
    return nodes_stage<T>(ctx, d_spheres, from_aabb, n, d_leaves, ctx->d_scalars + GB_SC_NLEAVES, w.leaf_deltas, w.flags,
                          w.leaf_boxes, d_nodes, d_root, st);
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

int launch_simple(const grace_b200_ctx* ctx, size_t n) { return grid_cap(ctx, n + 1, 256, 16); }

} // namespace

extern "C" {

int grace_b200_albvh_last_n_leaves(grace_b200_ctx* ctx, int* h_n_leaves, void* stream);

int grace_b200_deltas_euclid_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                float* d_deltas, void* stream)
{
    GB_REQUIRE(ctx && d_spheres4 && d_deltas, GRACE_B200_EINVAL, "NULL argument");
    deltas_euclid_kernel<<<launch_simple(ctx, n), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d_spheres4, n, d_deltas);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_deltas_sarea_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                               float* d_deltas, void* stream)
{
    GB_REQUIRE(ctx && d_spheres4 && d_deltas, GRACE_B200_EINVAL, "NULL argument");
    deltas_sarea_kernel<<<launch_simple(ctx, n), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d_spheres4, n, d_deltas);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_deltas_xor32(grace_b200_ctx* ctx, const uint32_t* d_keys, size_t n,
                            uint32_t* d_deltas, void* stream)
{
    GB_REQUIRE(ctx && d_keys && d_deltas, GRACE_B200_EINVAL, "NULL argument");
    deltas_xor_kernel<uint32_t><<<launch_simple(ctx, n), 256, 0, (cudaStream_t)stream>>>(d_keys, n, d_deltas);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

int grace_b200_deltas_xor64(grace_b200_ctx* ctx, const uint64_t* d_keys, size_t n,
                            uint64_t* d_deltas, void* stream)
{
    GB_REQUIRE(ctx && d_keys && d_deltas, GRACE_B200_EINVAL, "NULL argument");
    deltas_xor_kernel<uint64_t><<<launch_simple(ctx, n), 256, 0, (cudaStream_t)stream>>>(d_keys, n, d_deltas);
    GB_LAUNCH_CHECK();
    return GRACE_B200_OK;
}

static int albvh_build_any(grace_b200_ctx* ctx, const float* d_prims, bool from_aabb, size_t n,
                           const void* d_deltas, int delta_type, int max_per_leaf,
                           void* d_nodes, void* d_leaves, int* d_root, int* h_n_leaves, void* stream)
{
    GB_REQUIRE(ctx && d_prims && d_deltas && d_nodes && d_leaves && d_root, GRACE_B200_EINVAL,
               "NULL argument");
    GB_REQUIRE(max_per_leaf >= 1, GRACE_B200_EINVAL, "max_per_leaf must be >= 1");
    // albvh.cuh:795-799
    GB_REQUIRE(n > (size_t)max_per_leaf, GRACE_B200_EINVAL,
               "max_per_leaf must be less than the total number of primitives.");
    GB_REQUIRE(n < (1ull << 31) - 1024, GRACE_B200_ERANGE, "more than 2^31 primitives");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (delta_type == GRACE_B200_DELTA_F32)
        rc = build_typed<float>(ctx, (const float4*)d_prims, n, (const float*)d_deltas,
                                max_per_leaf, (int4*)d_nodes, (int4*)d_leaves, d_root, st, from_aabb);
    else if (delta_type == GRACE_B200_DELTA_U32)
        rc = build_typed<uint32_t>(ctx, (const float4*)d_prims, n, (const uint32_t*)d_deltas,
                                   max_per_leaf, (int4*)d_nodes, (int4*)d_leaves, d_root, st, from_aabb);
    else if (delta_type == GRACE_B200_DELTA_U64)
        rc = build_typed<uint64_t>(ctx, (const float4*)d_prims, n, (const uint64_t*)d_deltas,
                                   max_per_leaf, (int4*)d_nodes, (int4*)d_leaves, d_root, st, from_aabb);
    else
        return gb_set_error(GRACE_B200_EINVAL, "unknown delta_type %d", delta_type);
    if (rc) return rc;
    if (h_n_leaves) return grace_b200_albvh_last_n_leaves(ctx, h_n_leaves, stream);
    return GRACE_B200_OK;
}

int grace_b200_albvh_build_f4(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                              const void* d_deltas, int delta_type, int max_per_leaf,
                              void* d_nodes, void* d_leaves, int* d_root, int* h_n_leaves,
                              void* stream)
{
    return albvh_build_any(ctx, d_spheres4, false, n, d_deltas, delta_type, max_per_leaf, d_nodes, d_leaves,
                           d_root, h_n_leaves, stream);
}

int grace_b200_albvh_build_aabb(grace_b200_ctx* ctx, const float* d_aabbs8, size_t n,
                                const void* d_deltas, int delta_type, int max_per_leaf,
                                void* d_nodes, void* d_leaves, int* d_root, int* h_n_leaves,
                                void* stream)
{
    return albvh_build_any(ctx, d_aabbs8, true, n, d_deltas, delta_type, max_per_leaf, d_nodes, d_leaves,
                           d_root, h_n_leaves, stream);
}

int grace_b200_albvh_leaves(grace_b200_ctx* ctx, const void* d_deltas, int delta_type, size_t n, int max_per_leaf,
                            void* d_leaves, int* h_n_leaves, void* stream)
{
    GB_REQUIRE(ctx && d_deltas && d_leaves, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(max_per_leaf >= 1, GRACE_B200_EINVAL, "max_per_leaf must be >= 1");
    // albvh.cuh:795-799
    GB_REQUIRE(n > (size_t)max_per_leaf, GRACE_B200_EINVAL,
               "max_per_leaf must be less than the total number of primitives.");
    GB_REQUIRE(n < (1ull << 31) - 1024, GRACE_B200_ERANGE, "more than 2^31 primitives");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
#define GB_LEAVES_CASE(T)                                                                          \
    { BuildWs<T> w;                                                                                \
      if ((rc = build_workspace<T>(ctx, n, &w))) return rc;                                        \
      rc = leaves_stage<T>(ctx, n, (const T*)d_deltas, max_per_leaf, (int4*)d_leaves, w, st); }
    if (delta_type == GRACE_B200_DELTA_F32) GB_LEAVES_CASE(float)
    else if (delta_type == GRACE_B200_DELTA_U32) GB_LEAVES_CASE(uint32_t)
    else if (delta_type == GRACE_B200_DELTA_U64) GB_LEAVES_CASE(uint64_t)
    else return gb_set_error(GRACE_B200_EINVAL, "unknown delta_type %d", delta_type);
#undef GB_LEAVES_CASE
    if (rc) return rc;
    ctx->leaves_stage_n = n;
    ctx->leaves_stage_delta_type = delta_type;
    if (h_n_leaves) return grace_b200_albvh_last_n_leaves(ctx, h_n_leaves, stream);
    return GRACE_B200_OK;
}

static int albvh_nodes_any(grace_b200_ctx* ctx, const float* d_prims, bool from_aabb, const void* d_leaves,
                           size_t n_leaves, const void* d_leaf_deltas, int delta_type, void* d_nodes, int* d_root,
                           void* stream)
{
    GB_REQUIRE(ctx && d_prims && d_leaves && d_nodes && d_root, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_leaves >= 2 && n_leaves < (1ull << 31) - 1024, GRACE_B200_EINVAL, "a tree needs at least two leaves");
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_leaf_deltas) {
        // straight after grace_b200_albvh_leaves: its leaf-level deltas, zeroed arrival flags and leaf
        // count are still in the workspace / device scalars
        GB_REQUIRE(ctx->leaves_stage_n > 0 && ctx->leaves_stage_delta_type == delta_type, GRACE_B200_EINVAL,
                   "d_leaf_deltas is NULL but no grace_b200_albvh_leaves call with this delta type precedes");
        const size_t n = ctx->leaves_stage_n;
        ctx->leaves_stage_n = 0;
        int* d_nl = ctx->d_scalars + GB_SC_NLEAVES;
#define GB_NODES_CASE(T)                                                                                   \
        { BuildWs<T> w;                                                                                    \
          int rc = build_workspace<T>(ctx, n, &w);                                                         \
          if (rc) return rc;                                                                               \
          return nodes_stage<T>(ctx, (const float4*)d_prims, from_aabb, n_leaves, (const int4*)d_leaves, d_nl, w.leaf_deltas, \
                                w.flags, w.leaf_boxes, (int4*)d_nodes, d_root, st); }
        if (delta_type == GRACE_B200_DELTA_F32) GB_NODES_CASE(float)
        if (delta_type == GRACE_B200_DELTA_U32) GB_NODES_CASE(uint32_t)
        if (delta_type == GRACE_B200_DELTA_U64) GB_NODES_CASE(uint64_t)
#undef GB_NODES_CASE
        return gb_set_error(GRACE_B200_EINVAL, "unknown delta_type %d", delta_type);
    }
    const size_t flag_bytes = gb_align(n_leaves * sizeof(unsigned));
    unsigned* flags = (unsigned*)gb_workspace(ctx, flag_bytes + gb_align(2 * n_leaves * sizeof(float4)) + 256);
    if (!flags) return GRACE_B200_ENOMEM;
    float4* boxes = (float4*)((char*)flags + flag_bytes);
    GB_CUDA(cudaMemsetAsync(flags, 0, n_leaves * sizeof(unsigned), st));
    int* d_nleaves = ctx->d_scalars + GB_SC_NLEAVES;
    set_int_kernel<<<1, 1, 0, st>>>(d_nleaves, (int)n_leaves);
    GB_LAUNCH_CHECK();
    if (delta_type == GRACE_B200_DELTA_F32)
        return nodes_stage<float>(ctx, (const float4*)d_prims, from_aabb, n_leaves, (const int4*)d_leaves, d_nleaves,
                                  (const float*)d_leaf_deltas, flags, boxes, (int4*)d_nodes, d_root, st);
    if (delta_type == GRACE_B200_DELTA_U32)
        return nodes_stage<uint32_t>(ctx, (const float4*)d_prims, from_aabb, n_leaves, (const int4*)d_leaves, d_nleaves,
                                     (const uint32_t*)d_leaf_deltas, flags, boxes, (int4*)d_nodes, d_root, st);
    if (delta_type == GRACE_B200_DELTA_U64)
        return nodes_stage<uint64_t>(ctx, (const float4*)d_prims, from_aabb, n_leaves, (const int4*)d_leaves, d_nleaves,
                                     (const uint64_t*)d_leaf_deltas, flags, boxes, (int4*)d_nodes, d_root, st);
    return gb_set_error(GRACE_B200_EINVAL, "unknown delta_type %d", delta_type);
}

int grace_b200_albvh_nodes_f4(grace_b200_ctx* ctx, const float* d_spheres4, const void* d_leaves, size_t n_leaves,
                              const void* d_leaf_deltas, int delta_type, void* d_nodes, int* d_root, void* stream)
{
    return albvh_nodes_any(ctx, d_spheres4, false, d_leaves, n_leaves, d_leaf_deltas, delta_type, d_nodes, d_root, stream);
}

int grace_b200_albvh_nodes_aabb(grace_b200_ctx* ctx, const float* d_aabbs8, const void* d_leaves, size_t n_leaves,
                                const void* d_leaf_deltas, int delta_type, void* d_nodes, int* d_root, void* stream)
{
    return albvh_nodes_any(ctx, d_aabbs8, true, d_leaves, n_leaves, d_leaf_deltas, delta_type, d_nodes, d_root, stream);
}

int grace_b200_albvh_last_n_leaves(grace_b200_ctx* ctx, int* h_n_leaves, void* stream)
{
    GB_REQUIRE(ctx && h_n_leaves, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    GB_CUDA(cudaMemcpyAsync(ctx->h_pinned + GB_SC_NLEAVES, ctx->d_scalars + GB_SC_NLEAVES,
                            sizeof(int), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    *h_n_leaves = ctx->h_pinned[GB_SC_NLEAVES];
    return GRACE_B200_OK;
}

} // extern "C"
