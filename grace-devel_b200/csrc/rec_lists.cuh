// rec_lists.cuh -- one-pass hit lists, second half: where each unit's hits start, the copy order, the copy.
// (First half -- the recording traversal, pk_flush_rec / pk_rec_close -- is in trace_packet.cuh.)
#pragma once

// Traversal order of a ray's hits: the unit's own, then what was stolen from it, latest theft first
// (thefts take the BOTTOM of the stack, i.e. what the unit would have walked last), each stolen subtree
// recursively the same.  So a pre-order walk of a packet's theft tree with a running per-ray position
// gives every task the position of its first hit; it is stored in the second half of the task's record
// (the first half holds its own hit count per ray).  One warp per robbed packet, lane = ray.
constexpr int PK_RESOLVE_DEPTH = 128;
__global__ void __launch_bounds__(128) rec_resolve_kernel(const int2* __restrict__ roots, const int* __restrict__ root_own,
                                                         int* records, const int* __restrict__ offsets, int n_packets,
                                                         const int* __restrict__ created_ptr, int records_cap, int* overflow)
{
    const int created = min(__ldg(created_ptr), records_cap);
    __shared__ int stk[4][PK_RESOLVE_DEPTH];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int packet = blockIdx.x * 4 + w; packet < n_packets; packet += gridDim.x * 4) {
        int node = __ldg(&roots[packet].y);
        if (node < 0) continue;
        int cur = __ldg(offsets + packet * 32 + lane) + __ldg(root_own + packet * 32 + lane);
        int sp = 0;
        for (int guard = 0;; ++guard) {
            if (node < 0) {
                if (sp == 0) break;
                node = stk[w][--sp];
            }
            if (node >= created || guard > created) { if (lane == 0) atomicExch(overflow, 1); break; }     // malformed: two passes
            int* rec = records + (size_t)node * PK_DREC_WORDS;
            const int older = __ldcg(rec + PK_DR_OLDER), dons = __ldcg(rec + PK_DR_DONS), status = __ldcg(rec + PK_DR_STATUS);
            const int own = __ldcg(rec + PK_DR_ENTRIES + lane);
            if (status != 1) { if (lane == 0) atomicExch(overflow, 1); break; }
            rec[PK_DR_ENTRIES + 32 + lane] = cur;
            cur += own;
            if (older >= 0) {
                if (sp >= PK_RESOLVE_DEPTH) { if (lane == 0) atomicExch(overflow, 1); break; }
                __syncwarp();
                if (lane == 0) stk[w][sp] = older;
                ++sp;
                __syncwarp();
            }
            node = dons;
        }
        __syncwarp();
    }
}

// Copy order.  When a unit ends it takes one slot per GROUP of PK_COPY_G consecutive chunks of its own
// (pk_rec_close); this kernel files every chunk under (its unit's first slot + its number / PK_COPY_G).
// `order` was set to -1 beforehand: the last group of a unit may be short.
__global__ void __launch_bounds__(256) rec_order_kernel(const char* __restrict__ pool, const int* __restrict__ pool_ctr, int pool_cap,
                                                       const int2* __restrict__ roots, const int* __restrict__ records,
                                                       int* __restrict__ order)
{
    const int n_chunks = min(__ldg(pool_ctr), pool_cap);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
        const int4 h = __ldg((const int4*)(pool + (size_t)c * PK_RCH_BYTES));     // entries, unit, which, -
        const int start = h.y >= 0 ? __ldg(&roots[h.y].x) : __ldg(records + (size_t)(-1 - h.y) * PK_DREC_WORDS + PK_DR_HEAD);
        const int group = start + h.z / PK_COPY_G;
        if (group >= 0 && group < pool_cap) order[(size_t)group * PK_COPY_G + h.z % PK_COPY_G] = c;
    }
}

// The copy.  A warp takes a group: up to PK_COPY_G consecutive chunks of ONE unit, i.e. for every ray a run
// of consecutive hits of that ray (entry {integral, distance, index, k << 5 | lane} is hit k of ray `lane`
// within the unit).  Written entry by entry those runs arrive 3-5 hits (12-20 bytes) at a time, flush by
// flush, and partly written 32-byte sectors went to DRAM and back: 7.3 GB of traffic for 4.7 GB of payload.
// So the group is staged in shared memory (cp.async), an index ordered by (ray, k) is built -- per-ray
// count and first k by shared-memory atomics, a shuffle scan, one more pass -- and the arrays are written in
// that order: consecutive lanes, consecutive positions, whole sectors except at the ends of a run.
constexpr int RC_WARPS = 2;
constexpr int RC_MAXE = PK_COPY_G * PK_RCH_CAP;
static_assert(RC_MAXE <= 2048, "entry numbers are packed into 11 bits");
struct RcWarp {
    float4 raw[RC_MAXE];
    unsigned short idx[RC_MAXE + 8];
    int kmin[32], cnt[32], rowstart[32];
};

__global__ void __launch_bounds__(RC_WARPS * 32) rec_copy_kernel(const char* __restrict__ pool, const int* __restrict__ n_groups_ptr, int groups_cap,
                                                                const int* __restrict__ order, const int* __restrict__ records,
                                                                const int* __restrict__ offsets,
                                                                int* __restrict__ hit_idx, float* __restrict__ hit_integral, float* __restrict__ hit_dist)
{
    extern __shared__ __align__(16) unsigned char rc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RcWarp& W = ((RcWarp*)rc_smem)[warp];
    const int n_groups = min(__ldg(n_groups_ptr), groups_cap);
    // the caller's arrays: keep the lines in L2 for the ends of the runs and for the sort that usually follows
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (int g = blockIdx.x * RC_WARPS + warp; g < n_groups; g += gridDim.x * RC_WARPS) {
        int my_chunk = -1;
        if (lane < PK_COPY_G) my_chunk = __ldg(order + (size_t)g * PK_COPY_G + lane);
        int2 h = make_int2(0, 0);
        if (my_chunk >= 0) h = __ldg((const int2*)(pool + (size_t)my_chunk * PK_RCH_BYTES));      // entries, unit
        const int my_n = my_chunk >= 0 ? min(max(h.x, 0), PK_RCH_CAP) : 0;
        int my_e0 = 0, total = 0;          // where the chunk's entries start in the staging area
#pragma unroll
        for (int c = 0; c < PK_COPY_G; ++c) {
            const int n = __shfl_sync(0xffffffffu, my_n, c);
            if (lane == c) my_e0 = total;
            total += n;
        }
        const int unit = __shfl_sync(0xffffffffu, h.y, 0);
        if (total == 0 || __shfl_sync(0xffffffffu, my_chunk, 0) < 0) continue;
        const int base = unit >= 0 ? __ldg(offsets + (size_t)unit * 32 + lane)
                                   : __ldg(records + (size_t)(-1 - unit) * PK_DREC_WORDS + PK_DR_ENTRIES + 32 + lane);
        __syncwarp();                      // the previous group's reads of the staging area are done
#pragma unroll
        for (int c = 0; c < PK_COPY_G; ++c) {
            const int chunk = __shfl_sync(0xffffffffu, my_chunk, c);
            const int n = __shfl_sync(0xffffffffu, my_n, c), e0 = __shfl_sync(0xffffffffu, my_e0, c);
            if (chunk < 0) continue;
            const float4* e = (const float4*)(pool + (size_t)chunk * PK_RCH_BYTES + 32);
            for (int k = lane; k < n; k += 32) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(&W.raw[e0 + k]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(e + k) : "memory");
            }
        }
        W.kmin[lane] = 0x7fffffff; W.cnt[lane] = 0;
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // per ray: number of entries and the first k
        for (int e = lane; e < total; e += 32) {
            const unsigned code = (unsigned)__float_as_int(W.raw[e].w);
            atomicMin(&W.kmin[code & 31u], (int)(code >> 5));
            atomicAdd(&W.cnt[code & 31u], 1);
        }
        __syncwarp();
        const int c_l = W.cnt[lane], km = W.kmin[lane];
        int incl = c_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int rs = incl - c_l;
        W.rowstart[lane] = rs - km;        // slot of hit k of this ray = rowstart + k
        __syncwarp();
        for (int e = lane; e < total; e += 32) {
            const unsigned code = (unsigned)__float_as_int(W.raw[e].w);
            const int slot = W.rowstart[code & 31u] + (int)(code >> 5);
            W.idx[min(max(slot, 0), RC_MAXE - 1)] = (unsigned short)(e | ((code & 31u) << 11));
        }
        __syncwarp();
        const int posbase = base + km - rs;          // position of slot i of this ray = posbase + i
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            const unsigned v = i < total ? W.idx[i] : 0u;
            const int pb = __shfl_sync(0xffffffffu, posbase, v >> 11);
            if (i < total) {
                const float4 x = W.raw[v & 2047u];
                const int pos = pb + i;
                asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(hit_integral + pos), "f"(x.x), "l"(pol) : "memory");
                asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(hit_dist + pos), "f"(x.y), "l"(pol) : "memory");
                asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(hit_idx + pos), "f"(x.z), "l"(pol) : "memory");
            }
        }
    }
}
