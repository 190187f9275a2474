// trace_packet.cuh -- the production packet traversal kernel (included by trace.cu).
//
// Same packet structure as the reference (gpu::trace_kernel, cuda/kernels/bintree_trace.cuh:
// 52-197: 32 consecutive rays share one traversal, a child is entered if ANY lane's slab test
// reaches it) because rays of a packet overlap heavily -- measured on the 2^24-particle
// workload: a leaf visited by a packet is needed by ~21 of its 32 rays -- so node fetches and
// leaf staging are shared 32 ways.  What changes is where the cycles go:
//
//   * slab test against boxes padded by 64 ulp of the ray's coordinate scale, evaluated as
//     t = fma(plane, 1/d, -(o -/+ pad)/d).  Padding makes the visited set a superset of every
//     leaf holding a sphere that sphere_test() accepts (DESIGN.md "conservative slab test"),
//     so the hit set is exactly the brute-force set the reference's own test demands;
//   * per-lane hit masks travel with the traversal stack, so every leaf knows which rays can
//     touch it;
//   * memory-level parallelism: the walk first collects the next PK_BATCH leaves, fetches
//     their records in one round trip and streams all their spheres into shared memory with
//     cp.async in another, instead of three dependent round trips per leaf;
//   * at staging, spheres that no ray of the packet can hit (outside a conservative
//     cone/cylinder around the packet) are dropped and the rest compacted; the staging lane
//     precomputes h*h, 1/h (IEEE) and, when all 32 rays share one origin, c - o;
//   * dense leaves (many rays, few spheres): every lane tests every staged sphere;
//     sparse leaves (few rays, many spheres): the work is transposed -- lane j holds sphere
//     j and the few active rays are visited one by one;
//   * the on-hit work (sqrt, double-precision table lerp) is not done in the test loop,
//     where it would run on ~1 lane in 9: hits go to a per-lane FIFO in shared memory and
//     are evaluated in batches.  FIFO order keeps each ray's accumulation in ascending
//     primitive order, so sums are bit-identical to the reference's;
//   * load balance: heavy packets are as wide as light ones (max/mean hits per lane of the
//     heaviest packets: 1.04-1.3) but run 10-25x longer, so the work is split by SUBTREE, not
//     by ray, and inside the launch, by work stealing: a warp without a packet picks the
//     longest-running unit it can see and asks it for the bottom entry of its stack (the
//     subtree it would reach last, the largest it holds), which it then walks for all 32
//     rays.  Hit counts are order-free (atomic adds).  Column densities need each ray's terms
//     in ascending primitive order: a task appends its {W, 1/h^2} terms to a chain of chunks
//     and a second launch folds the chains in traversal order, one FFMA per term, exactly as
//     the unsplit traversal would have (pk_fold_chain).  Hit lists keep the older ray-subset
//     split over several launches (their write positions depend on the hits before them).
//
// All shared-memory addresses are compile-time offsets of one per-warp struct (the first
// version derived them from the runtime max_per_leaf and ptxas re-derived them inside the
// hot loops: 18 integer instructions per test iteration).
#pragma once

#ifdef PK_DEBUG_LB
__device__ unsigned long long pk_dbg[32];
__device__ __forceinline__ unsigned long long pk_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PK_DBG(x) x
#else
#define PK_DBG(x)
#endif

constexpr int PK_THREADS = 128;
constexpr int PK_WARPS = PK_THREADS / 32;
#ifndef PK_MIN_BLOCKS_V
#define PK_MIN_BLOCKS_V 6
#endif
constexpr int PK_MIN_BLOCKS = PK_MIN_BLOCKS_V;   // hit lists, PROF, WIDE: 65536 / (6 * 128) = 85 registers; the
                                                 // other modes ask for 7 CTAs (72 registers), see the kernel
#ifndef PK_STACK_V
#define PK_STACK_V 192
#endif
constexpr int PK_STACK = PK_STACK_V;  // reference STACK_SIZE is 64 (kernel_config.h:13); 192 entries x 8 B is
                                      // what still lets 7 column-density CTAs fit in 228 KB of shared memory
constexpr int PK_QD = 8;             // FIFO depth per lane
#ifndef PK_BATCH_V
#define PK_BATCH_V 4
#endif
constexpr int PK_BATCH = PK_BATCH_V; // leaves fetched together
#ifndef PK_ROOM_MIN_V
#define PK_ROOM_MIN_V 2
#endif
constexpr int PK_ROOM_MIN = PK_ROOM_MIN_V;   // flush when fewer than this many pushes are guaranteed to fit

template <int MODE, int M4>
struct PkWarp {
    static constexpr bool NEED_Q = (MODE == MODE_CUMULATIVE || MODE == MODE_FILL || MODE == MODE_REC);
    static constexpr bool NEED_I = (MODE == MODE_FILL || MODE == MODE_REC);
    float4 prims[M4];                          // staged leaf {x,y,z,h*h} or {c-o, h*h}
    float4 rays[64];                           // ray r: {dx,dy,dz,ox} {oy,oz,len,-}
    float4 raw[PK_BATCH * M4];                 // cp.async landing zone
    int2 stack[PK_STACK];                      // {node or leaf index, lane mask}
    float ir[NEED_Q ? M4 : 4];                 // 1/h of the staged spheres
    float2 q[NEED_Q ? PK_QD * 32 : 2];         // FIFO {b2, 1/h}, [slot][lane]
    unsigned char cells[NEED_Q ? PK_QD * 32 : 4];   // occupied FIFO cells, slot-major (flush work list)
    int idx[NEED_I ? M4 : 4];                  // primitive index of the staged spheres
    float2 q2[NEED_I ? PK_QD * 32 : 2];        // FIFO {distance, index bits}
    // term chain of a recording task (column densities, stolen subtrees): see pk_chain_append
    int ch_on, ch_cur, ch_head, ch_fail;       // recording?  current / first chunk, pool exhausted
    int ch_off, ch_nb;                         // bytes and blocks used in the current chunk (hit records: entries used, -)
    int ch_unit;                               // hit records: the unit the chunk belongs to (packet, or -1 - theft record)
    char* pool; int* pool_ctr; int pool_cap;
};

// Term chains.  A chunk is PK_CH_BYTES: a 32-byte header {int next; int blocks; ...}, a directory
// of PK_CH_NB block descriptors {lane mask, rows | start << 8} and the terms.  One block per FIFO
// flush: `rows` rows of one {W, 1/h^2} per ACTIVE lane (a lane with at least one term in this
// flush), lanes with fewer terms padded with {0, 0} -- fma(0, 0, x) = x -- so a term's address needs
// no prefix sum, and a task in which three of the 32 rays hit anything stores three columns (a
// fixed [row][32 lanes] layout took 6x the space on incoherent packets).  Each lane's terms are in
// emission (= ascending primitive) order.  The directory lets the fold know every address of a
// chunk after ONE dependent load.  Chunks come from a pool by atomic ticket.
constexpr int PK_CH_BYTES = 16384;
constexpr int PK_CH_NB = 64;
constexpr int PK_CH_TERMS = 32 + 8 * PK_CH_NB;     // byte offset of the first term

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p)
{
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

// cuda/functors/trace.cuh:183-186 + generic/interpolate.h:15-38 with 1/h precomputed (the
// same IEEE quotient the reference forms per hit).  table[i] = {T[i], T[i+1] - T[i]}: the
// reference forms the same double difference per hit.  t = x - i is exact in float
// (Sterbenz: i <= x < i + 1) so it equals the reference's (double)x - (double)i.
__device__ __forceinline__ float pk_lerp(float b2, float ir, const double2* table)
{
    const float x = __fmul_rn(__fmul_rn(__fsqrt_rn(b2), ir), 50.0f);
    const int i_raw = __float2int_rz(x);
    const int i = min(max(i_raw, 0), N_TABLE - 2);
    const float t = i_raw >= N_TABLE - 1 ? 1.0f : __fsub_rn(x, (float)i);
    const double2 e = table[i];
    return (float)__fma_rn((double)t, e.y, e.x);
}

#define NEED_I_(mode) ((mode) == MODE_FILL || (mode) == MODE_REC)     // hit lists: the FIFO also carries {distance, index}

// Everything a packet accumulates per lane, plus where hit lists go.
struct PkAcc {
    int count; float cum; int cursor;
    int qn;       // entries in this lane's FIFO
    int room;     // warp-uniform: pushes per lane guaranteed to fit before the next check
    int* hit_idx; float* hit_integral; float* hit_dist;
};

// Evaluate all FIFOs.  The on-hit work (sqrt, double-precision lerp) is spread evenly over
// the 32 lanes whatever the per-ray hit counts are: the occupied cells are listed slot-major
// (conflict-free), each lane evaluates every 32nd listed cell and writes {W, 1/h^2} back in
// place; then every lane folds ITS OWN cells in FIFO order -- ascending primitive order, one
// FFMA each, exactly the reference's accumulation (OnHit_sphere_cumulate is one fma in SASS).
// Real calls (not inlined): the kernel has ~10 call sites.
template <int MODE, int M4>
__device__ __forceinline__ void pk_flush_eval(PkWarp<MODE, M4>& W, int qn, int lane, unsigned lt, const double2* table)
{
    __syncwarp();       // entries may have been written by other lanes (transposed leaves)
    int total = 0;
#pragma unroll
    for (int j = 0; j < PK_QD; ++j) {
        const unsigned m = __ballot_sync(0xffffffffu, j < qn);
        if (j < qn) W.cells[total + __popc(m & lt)] = (unsigned char)(j * 32 + lane);
        total += __popc(m);
    }
    __syncwarp();
    for (int k = lane; k < total; k += 32) {
        const int c = W.cells[k];
        const float2 e = W.q[c];
        const float w = pk_lerp(e.x, e.y, table);
        const float ir2 = __fmul_rn(e.y, e.y);
        // per-hit value as stored (OnHit_sphere_individual): one FMUL
        W.q[c] = MODE == MODE_FILL ? make_float2(__fmul_rn(w, ir2), 0.f) : make_float2(w, ir2);
    }
    __syncwarp();
}


// Append the evaluated FIFO cells of all lanes to the unit's term chain as one block.  Out of line:
// pk_flush_cum is on the hot path of every packet and must not carry this function's registers.
template <int MODE, int M4>
__device__ __noinline__ void pk_chain_append(PkWarp<MODE, M4>& W, int qn, int lane, unsigned lt)
{
    constexpr int EB = 8;                             // bytes per term {W, 1/h^2}
    const unsigned mask = __ballot_sync(0xffffffffu, qn > 0);
    if (mask == 0u) return;
    const int rows = __reduce_max_sync(0xffffffffu, qn);
    const int ncols = __popc(mask), rank = __popc(mask & lt);
    const int size = EB * rows * ncols;
    int cur = W.ch_cur, off = W.ch_off, nb = W.ch_nb;
    if (cur < 0 || nb == PK_CH_NB || off + size > PK_CH_BYTES) {
        if (W.ch_fail) return;                   // pool exhausted earlier: the unit is about to abort
        int id = -1;
        if (lane == 0) { id = atomicAdd(W.pool_ctr, 1); if (id >= W.pool_cap) id = -1; }
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id < 0) { if (lane == 0) W.ch_fail = 1; __syncwarp(); return; }
        if (lane == 0) {
            if (cur >= 0) *(int2*)(W.pool + (size_t)cur * PK_CH_BYTES) = make_int2(id, nb);      // close and link
            else W.ch_head = id;
            W.ch_cur = id;
        }
        cur = id;
        off = PK_CH_TERMS;
        nb = 0;
    }
    char* ch = W.pool + (size_t)cur * PK_CH_BYTES;
    if (lane == 0) ((int2*)(ch + 32))[nb] = make_int2((int)mask, rows | (off << 8));
    if (qn > 0) {
        float2* t = (float2*)(ch + off) + rank;
#pragma unroll
        for (int j = 0; j < PK_QD; ++j)
            if (j < rows) t[j * ncols] = j < qn ? W.q[j * 32 + lane] : make_float2(0.f, 0.f);
    }
    __syncwarp();
    if (lane == 0) { W.ch_off = off + size; W.ch_nb = nb + 1; }
}

// Write the header of the last chunk (the unit is finished).
template <int MODE, int M4>
__device__ __forceinline__ void pk_chain_close(PkWarp<MODE, M4>& W, int lane)
{
    __syncwarp();
    const int cur = W.ch_cur;
    if (cur >= 0 && lane == 0) *(int2*)(W.pool + (size_t)cur * PK_CH_BYTES) = make_int2(-1, W.ch_nb);
}

// Fold a finished chain into `cum`: each lane adds ITS terms in the order they were emitted,
// one FFMA per term -- the arithmetic of pk_flush_cum, hence of the unsplit traversal.  Latency
// is what matters here (a heavy ray's chain is walked by one warp): one dependent load per chunk
// (header + directory), the next chunk prefetched into L2 as soon as its id is known, and the
// terms of block b + 1 requested before those of block b are added.  The chains were written by
// an earlier launch, so the read-only path may cache them.
__device__ __forceinline__ void pk_fold_load(const char* ch, int2 d, int lane, unsigned lt, float2 (&e)[PK_QD], int& rows)
{
    const unsigned mask = (unsigned)d.x;
    rows = ((mask >> lane) & 1u) ? (d.y & 0xff) : 0;
    const int ncols = __popc(mask);
    const float2* t = (const float2*)(ch + (d.y >> 8)) + __popc(mask & lt);
#pragma unroll
    for (int j = 0; j < PK_QD; ++j) if (j < rows) e[j] = __ldg(t + j * ncols);
}

constexpr int PK_FOLD_G = 8;      // blocks whose terms are requested together (the fold launch's kernel may use 255 registers)
__device__ __noinline__ float pk_fold_chain(const char* pool, int head, float cum, int lane, unsigned lt)
{
    if (head < 0) return cum;
    const char* ch = pool + (size_t)head * PK_CH_BYTES;
    int2 h = __ldg((const int2*)ch);             // next, blocks
    int2 d0 = __ldg((const int2*)(ch + 32) + lane), d1 = __ldg((const int2*)(ch + 32) + 32 + lane);
    for (;;) {
        // the next chunk's header and directory are requested now, its terms prefetched into L2
        int2 nh = make_int2(-1, 0), nd0 = make_int2(0, 0), nd1 = make_int2(0, 0);
        const char* nch = nullptr;
        if (h.x >= 0) {
            nch = pool + (size_t)h.x * PK_CH_BYTES;
            nh = __ldg((const int2*)nch);
            nd0 = __ldg((const int2*)(nch + 32) + lane);
            nd1 = __ldg((const int2*)(nch + 32) + 32 + lane);
#pragma unroll
            for (int k = 0; k < PK_CH_BYTES / 4096; ++k) asm volatile("prefetch.global.L2 [%0];" :: "l"(nch + lane * 128 + k * 4096));
        }
        const int nb = h.y;
        PK_DBG(if (lane == 0) { atomicAdd(&pk_dbg[20], 1ull); atomicAdd(&pk_dbg[21], (unsigned long long)nb); })
        for (int b0 = 0; b0 < nb; b0 += PK_FOLD_G) {
            const int2 dd = b0 < 32 ? d0 : d1;       // a group of 8 lies in one half of the directory
            float2 e[PK_FOLD_G][PK_QD];
            int rows[PK_FOLD_G];
#pragma unroll
            for (int g = 0; g < PK_FOLD_G; ++g) {
                const int2 d = make_int2(__shfl_sync(0xffffffffu, dd.x, (b0 + g) & 31), __shfl_sync(0xffffffffu, dd.y, (b0 + g) & 31));
                rows[g] = 0;
                if (b0 + g < nb) pk_fold_load(ch, d, lane, lt, e[g], rows[g]);
            }
#pragma unroll
            for (int g = 0; g < PK_FOLD_G; ++g)
#pragma unroll
                for (int j = 0; j < PK_QD; ++j) if (j < rows[g]) cum = __fmaf_rn(e[g][j].x, e[g][j].y, cum);
        }
        if (h.x < 0) break;
        ch = nch; h = nh; d0 = nd0; d1 = nd1;
    }
    return cum;
}

// ---- one-pass hit lists (MODE_REC) ----
// The counting traversal also records every hit, in chunks of PK_RCH_BYTES taken from a pool by atomic
// ticket: a 32-byte header {entries, unit, which of the unit's chunks, -} and up to PK_RCH_CAP entries
//     {integral, distance, primitive index, (k << 5) | lane}
// where k says that this is the k-th hit of ray `lane` WITHIN THE UNIT (a packet, or a stolen subtree of
// one).  Entries are self-describing, so a chunk can be copied to the caller's arrays by any warp once
// the position of its unit's first hit is known for each ray (rec_resolve_kernel, rec_lists.cuh); no chain has to
// be walked.  A flush lists the occupied FIFO cells ray-major exactly like pk_flush_fill, so the copy's
// stores land on a few contiguous stretches; it may straddle two chunks, nothing is padded.
constexpr int PK_RCH_BYTES = 8192;
constexpr int PK_RCH_CAP = (PK_RCH_BYTES - 32) / 16;
constexpr int PK_COPY_G = 4;                  // chunks of one unit that are copied together (rec_lists.cuh)
constexpr int PK_RCH_MAXK = 1 << 26;          // hits of one ray within one unit that (k << 5 | lane) can hold

template <int MODE, int M4>
__device__ __noinline__ int pk_flush_rec(PkWarp<MODE, M4>& W, int qn, int count, int lane, const double2* table)
{
    __syncwarp();       // entries may have been written by other lanes (transposed leaves)
    int incl = qn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int first = incl - qn;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return count;
    if (W.ch_fail) return count + qn;             // pool exhausted earlier: the counts stay right, the lists are dropped
    if (__any_sync(0xffffffffu, count + qn >= PK_RCH_MAXK)) { if (lane == 0) W.ch_fail = 1; __syncwarp(); return count + qn; }
#pragma unroll
    for (int j = 0; j < PK_QD; ++j)
        if (j < qn) W.cells[first + j] = (unsigned char)(j * 32 + lane);
    const int cur = W.ch_cur, used = W.ch_off;
    const int room = cur >= 0 ? PK_RCH_CAP - used : 0;
    int nxt = -1;
    if (total > room) {
        if (lane == 0) { nxt = atomicAdd(W.pool_ctr, 1); if (nxt >= W.pool_cap) nxt = -1; }
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        if (nxt < 0) { if (lane == 0) W.ch_fail = 1; __syncwarp(); return count + qn; }
    }
    __syncwarp();
    // cell k goes to d0[k] while the current chunk has room, to d1[k] in the new one
    float4* d0 = (float4*)(W.pool + (size_t)max(cur, 0) * PK_RCH_BYTES + 32) + used;
    float4* d1 = (float4*)(W.pool + (size_t)max(nxt, 0) * PK_RCH_BYTES + 32) - room;
    for (int base = 0; base < total; base += 32) {
        const int k = base + lane;
        const bool on = k < total;
        const int c = on ? W.cells[k] : 0;
        const int cnt = __shfl_sync(0xffffffffu, count, c & 31);      // the owning ray's hits so far
        if (on) {
            const float2 e = W.q[c], e2 = W.q2[c];
            const float w = pk_lerp(e.x, e.y, table);
            const float val = __fmul_rn(w, __fmul_rn(e.y, e.y));       // OnHit_sphere_individual: one FMUL
            const int code = ((cnt + (c >> 5)) << 5) | (c & 31);
            (k < room ? d0 : d1)[k] = make_float4(val, e2.x, e2.y, __int_as_float(code));
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (nxt >= 0) {
            const int nb = W.ch_nb;
            if (cur >= 0) *(int4*)(W.pool + (size_t)cur * PK_RCH_BYTES) = make_int4(PK_RCH_CAP, W.ch_unit, nb - 1, 0);     // full
            W.ch_cur = nxt; W.ch_off = total - room; W.ch_nb = nb + 1;
        } else W.ch_off = used + total;
    }
    __syncwarp();
    return count + qn;
}

// The unit is finished: header of its last chunk.  Returns the first of the consecutive slots of the copy
// order that the unit takes, one per group of PK_COPY_G consecutive chunks (a group is copied by one warp:
// for every ray it holds a run of consecutive hits of that ray).
template <int MODE, int M4>
__device__ __forceinline__ int pk_rec_close(PkWarp<MODE, M4>& W, int* slot_ctr, int lane)
{
    __syncwarp();
    const int cur = W.ch_cur;
    int start = 0;
    if (cur >= 0 && lane == 0) {
        const int nb = W.ch_nb;
        *(int4*)(W.pool + (size_t)cur * PK_RCH_BYTES) = make_int4(W.ch_off, W.ch_unit, nb - 1, 0);
        start = atomicAdd(slot_ctr, (nb + PK_COPY_G - 1) / PK_COPY_G);
    }
    return start;      // lane 0's is the one
}

template <int MODE, int M4>
__device__ __noinline__ float pk_flush_cum(PkWarp<MODE, M4>& W, int qn, float cum, int lane, unsigned lt, const double2* table)
{
#ifndef PK_NO_OWN_FLUSH
    // Well-filled FIFOs: every lane evaluates and adds its own cells -- no work list, no write-back.  The list
    // (cells spread evenly over the lanes) only pays when few lanes hold most of the cells; instruction counts
    // of the two paths from the SASS: ~33 per row here, ~116 + 38 per 32 listed cells there.  Same arithmetic,
    // same order per ray.
    if (!W.ch_on) {
        const int rows = __reduce_max_sync(0xffffffffu, qn), total = __reduce_add_sync(0xffffffffu, qn);
        if (rows * 33 <= 116 + ((total + 31) >> 5) * 38) {
            __syncwarp();       // entries may have been written by other lanes (transposed leaves)
#pragma unroll
            for (int j = 0; j < PK_QD; ++j) {
                if (j < rows) {
                    if (j < qn) {
                        const float2 e = W.q[j * 32 + lane];
                        cum = __fmaf_rn(pk_lerp(e.x, e.y, table), __fmul_rn(e.y, e.y), cum);
                    }
                }
            }
            __syncwarp();
            return cum;
        }
    }
#endif
    pk_flush_eval<MODE, M4>(W, qn, lane, lt, table);
    if (W.ch_on) {            // recording task: the terms go to the chain, the fold launch adds them
        pk_chain_append<MODE, M4>(W, qn, lane, lt);
        __syncwarp();
        return cum;
    }
#pragma unroll
    for (int j = 0; j < PK_QD; ++j)
        if (j < qn) { const float2 e = W.q[j * 32 + lane]; cum = __fmaf_rn(e.x, e.y, cum); }
    __syncwarp();
    return cum;
}

// Hit lists: evaluate and store in one go.  The occupied cells are listed RAY-major (a shuffle
// scan of the fill levels gives every lane the start of its run), so the 32 cells a step
// handles belong to a few rays and land on a few contiguous stretches of each output array:
// ~6 sectors per store instruction where one lane per ray wrote 32 (the fill pass was bound by
// those scattered 4-byte stores).  Values and positions are exactly those of the per-lane loop.
template <int MODE, int M4>
__device__ __noinline__ int pk_flush_fill(PkWarp<MODE, M4>& W, int qn, int cursor, int lane, unsigned lt, const double2* table,
                                          int* hit_idx, float* hit_integral, float* hit_dist)
{
    (void)lt;
    __syncwarp();       // entries may have been written by other lanes (transposed leaves)
    int incl = qn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int first = incl - qn;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int j = 0; j < PK_QD; ++j)
        if (j < qn) W.cells[first + j] = (unsigned char)(j * 32 + lane);
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int k = base + lane;
        const bool on = k < total;
        const int c = on ? W.cells[k] : 0;
        const int cur = __shfl_sync(0xffffffffu, cursor, c & 31);     // the owning ray's write position
        if (on) {
            const float2 e = W.q[c], e2 = W.q2[c];
            const float w = pk_lerp(e.x, e.y, table);
            const int pos = cur + (c >> 5);
            hit_idx[pos] = __float_as_int(e2.y);
            hit_integral[pos] = __fmul_rn(w, __fmul_rn(e.y, e.y));     // OnHit_sphere_individual: one FMUL
            hit_dist[pos] = e2.x;
        }
    }
    __syncwarp();
    return cursor + qn;
}

template <int MODE, int M4>
__device__ __forceinline__ void pk_flush(PkWarp<MODE, M4>& W, PkAcc& A, int lane, unsigned lt, const double2* table)
{
    if (MODE == MODE_CUMULATIVE) A.cum = pk_flush_cum<MODE, M4>(W, A.qn, A.cum, lane, lt, table);
    if (MODE == MODE_FILL)
        A.cursor = pk_flush_fill<MODE, M4>(W, A.qn, A.cursor, lane, lt, table, A.hit_idx, A.hit_integral, A.hit_dist);
    if (MODE == MODE_REC) {       // one-pass hit lists: evaluate, count, record (positions are assigned later)
        A.count = pk_flush_rec<MODE, M4>(W, A.qn, A.count, lane, table);
    }
    A.qn = 0;
    A.room = PK_QD;
}

// Re-derive how many pushes per lane are guaranteed to fit; flush when that is too few.
template <int MODE, int M4>
__device__ __forceinline__ void pk_room(PkWarp<MODE, M4>& W, PkAcc& A, int lane, unsigned lt, const double2* table)
{
    A.room = PK_QD - __reduce_max_sync(0xffffffffu, A.qn);
    if (A.room < PK_ROOM_MIN) pk_flush<MODE, M4>(W, A, lane, lt, table);
}

// ---- conservative bound of a whole packet --------------------------------------------
// Every point of every ray lies within r0 + max(0, t - t0min) * tan(theta) of the axis line
// (oc, a), t being the axial coordinate.  A sphere farther than that (+ h + slack) from the
// axis, or entirely behind every origin, cannot be hit by any ray of the packet.
struct PacketBound {
    float ax, ay, az;       // axis direction (unit)
    float ocx, ocy, ocz;    // a point on the axis (lane 0's origin)
    float r0, tan_t, t0min, tfar;
    bool enabled;
};

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ PacketBound packet_bound(const grace_b200_ray& ray)
{
    PacketBound B;
    const float sx = warp_sum(ray.dx), sy = warp_sum(ray.dy), sz = warp_sum(ray.dz);
    const float n2 = sx * sx + sy * sy + sz * sz;
    const float inv = rsqrtf(fmaxf(n2, 1e-30f));
    B.ax = sx * inv; B.ay = sy * inv; B.az = sz * inv;
    B.ocx = __shfl_sync(0xffffffffu, ray.ox, 0);
    B.ocy = __shfl_sync(0xffffffffu, ray.oy, 0);
    B.ocz = __shfl_sync(0xffffffffu, ray.oz, 0);
    const float wx = ray.ox - B.ocx, wy = ray.oy - B.ocy, wz = ray.oz - B.ocz;
    const float t0 = wx * B.ax + wy * B.ay + wz * B.az;
    const float qx = wx - t0 * B.ax, qy = wy - t0 * B.ay, qz = wz - t0 * B.az;
    const float perp = sqrtf(qx * qx + qy * qy + qz * qz);
    const float cosl = ray.dx * B.ax + ray.dy * B.ay + ray.dz * B.az;
    // sine from the perpendicular component, not from 1 - cos^2: a cosine that rounds to 1 still
    // allows an angle of 3.5e-4 rad
    const float sx_ = ray.dx - cosl * B.ax, sy_ = ray.dy - cosl * B.ay, sz_ = ray.dz - cosl * B.az;
    const float sinl = sqrtf(sx_ * sx_ + sy_ * sy_ + sz_ * sz_);
    const float cmin = warp_min(cosl);
    const float smax = warp_max(sinl);
    B.r0 = warp_max(perp) * 1.0001f;
    B.t0min = warp_min(t0);
    B.tfar = warp_max(t0 + fabsf(ray.length));      // no ray point lies beyond this axial coordinate
    // NaNs (degenerate rays) make the comparisons false -> culling disabled
    B.enabled = (n2 > 1e-12f) && (cmin > 0.5f) && (B.r0 < 1e30f) && (B.tfar < 1e30f);
    B.tan_t = smax / fminf(cmin, 1.0f) * 1.001f + 1e-6f;
    return B;
}

__device__ __forceinline__ bool packet_may_hit(const PacketBound& B, const float4 s)
{
    const float px = s.x - B.ocx, py = s.y - B.ocy, pz = s.z - B.ocz;
    const float t = px * B.ax + py * B.ay + pz * B.az;
    const float qx = px - t * B.ax, qy = py - t * B.ay, qz = pz - t * B.az;
    const float d2 = qx * qx + qy * qy + qz * qz;
    const float scale = fabsf(px) + fabsf(py) + fabsf(pz) + B.r0 + s.w;
    const float slack = 2e-5f * scale;
    const float R = B.r0 + fmaxf(0.0f, t + s.w - B.t0min) * B.tan_t + s.w + slack;
    const bool outside = d2 > R * R;
    const bool behind = (t + s.w + slack) < B.t0min;
    return !(outside || behind);
}

__device__ __forceinline__ float pk_finite_rcp(float d)
{
    const float big = 18446744073709551616.0f;      // 2^64
    const float r = __fdiv_rn(1.0f, d);
    return fabsf(r) <= big ? r : copysignf(big, d);  // inf and NaN (d = 0, -0, NaN) take the clamp
}

// Padded slab test of one ray against one box: t = fma(plane, 1/d, -(o -/+ pad)/d).
struct PkSlab { float ix, iy, iz, cbx, ctx, cby, cty, cbz, ctz, len; };
__device__ __forceinline__ bool pk_slab(const PkSlab& S, int bx, int tx, int by, int ty, int bz, int tz)
{
    const float a0 = fmaf(__int_as_float(bx), S.ix, S.cbx), a1 = fmaf(__int_as_float(tx), S.ix, S.ctx);
    const float b0 = fmaf(__int_as_float(by), S.iy, S.cby), b1 = fmaf(__int_as_float(ty), S.iy, S.cty);
    const float c0 = fmaf(__int_as_float(bz), S.iz, S.cbz), c1 = fmaf(__int_as_float(tz), S.iz, S.ctz);
    const float tmin = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), 0.0f));
    const float tmax = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), S.len));
    return !(tmax < tmin);
}

// One sphere against one ray: generic/intersect.h:16-48 in the SASS-verified contraction.
// COMMON: the staged xyz already hold c - o.  s.w holds h*h.
template <bool COMMON>
__device__ __forceinline__ bool pk_test(const float4 s, float ox, float oy, float oz,
                                        float dx, float dy, float dz, float len,
                                        float& b2, float& dot)
{
    const float px = COMMON ? s.x : __fsub_rn(s.x, ox);
    const float py = COMMON ? s.y : __fsub_rn(s.y, oy);
    const float pz = COMMON ? s.z : __fsub_rn(s.z, oz);
    dot = __fmul_rn(py, dy);
    dot = __fmaf_rn(px, dx, dot);
    dot = __fmaf_rn(pz, dz, dot);
    const float bx = __fmaf_rn(-dx, dot, px);
    const float by = __fmaf_rn(-dy, dot, py);
    const float bz = __fmaf_rn(-dz, dot, pz);
    b2 = __fmul_rn(by, by);
    b2 = __fmaf_rn(bx, bx, b2);
    b2 = __fmaf_rn(bz, bz, b2);
    return !(b2 >= s.w) && !(dot < 0.0f) && !(dot >= len);
}

// Dense leaf: every lane tests every staged sphere against its own ray.  Hits go to the
// lane's FIFO; the loop body has no vote and no branch: it runs for `room` spheres, the
// number of pushes every lane's FIFO is known to have space for, and only then looks at the
// fill levels again.
#ifndef PK_SPARSE_NUM
#define PK_SPARSE_NUM 3
#define PK_SPARSE_DEN 2
#endif
#ifndef PK_PAIRS_OK
#define PK_PAIRS_OK(mode) ((mode) == MODE_CUMULATIVE)      // hit lists carry twice the per-hit state: no gain there
#endif
template <int MODE, int M4, bool COMMON>
__device__ __forceinline__ void pk_leaf_dense(PkWarp<MODE, M4>& W, int n_kept, const grace_b200_ray& ray,
                                              int lane, unsigned lt, bool lane_on, PkAcc& A, const double2* table)
{
    if (MODE == MODE_COUNT) {
#pragma unroll 4
        for (int i = 0; i < n_kept; ++i) {
            float b2, dot;
            A.count += (int)(pk_test<COMMON>(W.prims[i], ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz,
                                             ray.length, b2, dot) && lane_on);
        }
        return;
    }
    int i = 0;
    while (i < n_kept) {
        if (A.room == 0) pk_room<MODE, M4>(W, A, lane, lt, table);
        const int n = min(n_kept - i, A.room);
        A.room -= n;
        const int end = i + n;
        int qn = A.qn;
        // two spheres per step: both tests are computed before either push, so their dependent
        // FMA chains overlap (the pushes' shared-memory stores would otherwise order them)
        for (; PK_PAIRS_OK(MODE) && i + 1 < end; i += 2) {
            const float4 s0 = W.prims[i], s1 = W.prims[i + 1];
            float b20, dot0, b21, dot1;
            const bool hit0 = pk_test<COMMON>(s0, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, ray.length, b20, dot0) && lane_on;
            const bool hit1 = pk_test<COMMON>(s1, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, ray.length, b21, dot1) && lane_on;
            if (hit0) {
                W.q[qn * 32 + lane] = make_float2(b20, W.ir[i]);
                if (NEED_I_(MODE)) W.q2[qn * 32 + lane] = make_float2(dot0, __int_as_float(W.idx[i]));
                ++qn;
            }
            if (hit1) {
                W.q[qn * 32 + lane] = make_float2(b21, W.ir[i + 1]);
                if (NEED_I_(MODE)) W.q2[qn * 32 + lane] = make_float2(dot1, __int_as_float(W.idx[i + 1]));
                ++qn;
            }
        }
        for (; i < end; ++i) {
            const float4 s = W.prims[i];
            float b2, dot;
            const bool hit = pk_test<COMMON>(s, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, ray.length, b2, dot) && lane_on;
            if (hit) {
                W.q[qn * 32 + lane] = make_float2(b2, W.ir[i]);
                if (NEED_I_(MODE)) W.q2[qn * 32 + lane] = make_float2(dot, __int_as_float(W.idx[i]));
                ++qn;
            }
        }
        A.qn = qn;
    }
}

// Sparse leaf: lane j holds staged sphere j; the rays in `mask` are visited one by one
// (shared-memory broadcast) and hits go straight into the owning ray's FIFO, in ascending
// sphere order.
template <int MODE, int M4, bool COMMON>
__device__ __forceinline__ void pk_leaf_sparse(PkWarp<MODE, M4>& W, unsigned mask, int n_kept, int lane,
                                               unsigned lt, PkAcc& A, const double2* table)
{
    for (int base = 0; base < n_kept; base += 32) {
        const int j = base + lane;
        const bool have = j < n_kept;
        const float4 s = have ? W.prims[j] : make_float4(0.f, 0.f, 0.f, -1.f);
        float ir = 0.f;
        int pidx = 0;
        if (MODE != MODE_COUNT && have) ir = W.ir[j];
        if (NEED_I_(MODE) && have) pidx = W.idx[j];
        unsigned m = mask;
        while (m) {
            const int r = __ffs(m) - 1;
            m &= m - 1;
            const float4 ra = W.rays[2 * r];
            const float4 rb = W.rays[2 * r + 1];
            float b2, dot;
            const bool hit = pk_test<COMMON>(s, ra.w, rb.x, rb.y, ra.x, ra.y, ra.z, rb.z, b2, dot) && have;
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hm == 0u) continue;
            const int nh = __popc(hm);
            if (MODE == MODE_COUNT) {
                if (lane == r) A.count += nh;
            } else {
                const int rank = __popc(hm & lt);
                int qn_r = __shfl_sync(0xffffffffu, A.qn, r);
                int done = 0;
                while (done < nh) {
                    if (qn_r == PK_QD) { pk_flush<MODE, M4>(W, A, lane, lt, table); qn_r = 0; }
                    const int take = min(PK_QD - qn_r, nh - done);
                    if (hit && rank >= done && rank < done + take) {
                        W.q[(qn_r + rank - done) * 32 + r] = make_float2(b2, ir);
                        if (NEED_I_(MODE)) W.q2[(qn_r + rank - done) * 32 + r] = make_float2(dot, __int_as_float(pidx));
                    }
                    if (lane == r) A.qn += take;
                    qn_r += take;
                    done += take;
                    A.room = min(A.room, PK_QD - qn_r);
                }
            }
        }
    }
}

// ---- load balancing -------------------------------------------------------------------------
// Hit counts and column densities (SUB = true): work stealing inside the launch, thief-directed.
// Every warp of the (persistent, fully resident) grid has a slot word `state` (0 = nothing to
// steal, s > 0 = the steps its unit has run + 1, refreshed every PK_CHECK_EVERY steps, s < 0 = a
// thief's request) and an answer cell `resp`.  A warp that finds no packet left becomes a thief: it samples a few hundred
// slots, picks the unit that has run longest (cost is heavy-tailed: the longest-running unit is
// the best guess for the one with most left) and turns that warp's slot word into its request with
// one compare-and-swap.  The victim sees the request at its next check and hands over the BOTTOM
// entry of its stack -- the subtree it would reach last, the largest it holds -- as a task for all
// rays of the unit, or answers "nothing" if it holds no subtree worth a task.  There is no shared
// queue and no counter everybody fights over (a first version with a global task queue lost 99 %
// of its donation attempts to compare-and-swap races and starved exactly the heavy packets).
// Tasks are stolen from in the same way.  Results:
//   counts:  order-free -- tasks and robbed packets atomicAdd into the (zeroed) output;
//   column densities: a ray's terms must be added in ascending primitive order, i.e. a unit's own
//     terms first, then the subtrees stolen from it from the LATEST to the first (later thefts
//     take entries from higher up the stack), recursively.  Packets add their own terms in
//     registers as always; tasks write theirs to chunk chains; a robbed packet leaves its sum in a
//     root slot and a second launch ("fold", one warp per root) walks the theft tree depth-first,
//     one cursor per nesting level on the ordinary traversal stack.  A task that could not get a
//     chunk (pool exhausted) is marked aborted and its subtree is walked by the fold unit itself,
//     in place -- slower, never wrong.
// Hit lists (SUB = false), over several launches: every hit's write position depends on the
// number of hits before it, so an over-budget unit is suspended once no unclaimed unit is left
// and resumed as tasks over disjoint subsets of 8, 2, 1 of its RAYS, each with the whole stack
// and its own cursors.
constexpr int PK_REC_WORDS = 8 + 2 * PK_STACK + 3 * 32;   // hit lists: header | stack | cum, count, cursor
// A theft: {packet, ray subset, number of stack entries, - | older theft from the same unit, then the
// task's results: first chunk of its term chain, latest theft from IT, status 0/1/2 | the entries}
constexpr int PK_DON_MAX = 32;                      // stack entries per theft: the bottom ones up to the first big subtree
constexpr int PK_DREC_WORDS = 8 + 2 * PK_DON_MAX;
enum { PK_DR_PACKET = 0, PK_DR_SUBSET = 1, PK_DR_N = 2, PK_DR_OLDER = 4, PK_DR_HEAD = 5, PK_DR_DONS = 6, PK_DR_STATUS = 7, PK_DR_ENTRIES = 8 };
#ifndef PK_CHECK_EVERY_V
#define PK_CHECK_EVERY_V 8
#endif
constexpr int PK_CHECK_EVERY = PK_CHECK_EVERY_V;    // steps between looks at the mailbox
#ifndef PK_DON_MINR_V
#define PK_DON_MINR_V 64
#endif
constexpr int PK_DON_MINR = PK_DON_MINR_V;          // smallest subtree (leaves spanned) worth a task
constexpr int PK_SAMPLE = 4;                        // 128-byte lines of slot words a thief samples per scan
constexpr int PK_ADV = 256;                         // notice board: slots of long-running units (hints, may be stale)
#ifndef PK_ADV_AGE_V
#define PK_ADV_AGE_V 256
#endif
constexpr int PK_ADV_AGE = PK_ADV_AGE_V;                     // steps after which a unit puts itself on the board
constexpr int PK_SPIN_LIMIT = 1 << 20;              // polls (up to ~4 us apart) before a waiting warp gives up (error 3)

enum { PK_KIND_PACKETS = 0, PK_KIND_TASKS = 1, PK_KIND_FOLD = 2 };
enum { PK_LB_FINISHED = 0, PK_LB_CREATED = 1, PK_LB_OVERFLOW = 7 };     // units finished; tasks created (= theft records); hit pool ran dry

struct PkTasks {
    int kind;
    int* records;             // theft records (PK_DREC_WORDS) or, hit lists, suspended traversals (PK_REC_WORDS)
    int* n_records;           // hit lists only
    int tasks_cap, records_cap;
    int budget;               // steps (inner nodes + leaves) before a unit may be robbed / suspended
    int eager;                // 1: any subtree is worth a task, suspension at `budget` regardless (tests)
    // ---- work stealing (counts, column densities) ----
    int* state; int* resp; int n_slots;       // one word per resident warp, see pk_serve
    int* adv;                                 // PK_ADV hints
    int* lb;                       // PK_LB_*
    char* pool; int* pool_ctr; int pool_cap;
    int2* roots; float* root_cum; int* n_roots;    // robbed packets: {packet, latest theft}, their own sums
    int* order;                    // hit records: chunks in copy order
    // ---- ray-subset rounds (hit lists) ----
    const int2* tasks_in;     // PK_KIND_TASKS: {record, ray subset}
    const int* n_tasks_in;
    int2* tasks_out;          // NULL: run to completion
    int* n_tasks_out;
    int child_width;          // lanes per child task
    // running statistics of the units that completed unsplit -- ray-subset tasks lose SIMD width,
    // so a unit is only split when it is much heavier than the typical one
    unsigned long long* sum_steps;
    int* n_done;
};
#ifndef PK_AVG_FACTOR_X4_V
#define PK_AVG_FACTOR_X4_V 8
#endif
constexpr int PK_AVG_FACTOR_X4 = PK_AVG_FACTOR_X4_V;   // split units heavier than FACTOR/4 x the mean

// Claim `n` consecutive slots of a bounded pool; -1 if they do not fit.
__device__ __forceinline__ int pk_reserve(int* counter, int n, int cap)
{
    int cur = *(volatile int*)counter;
    for (;;) {
        if (cur + n > cap) return -1;
        const int seen = atomicCAS(counter, cur, cur + n);
        if (seen == cur) return cur;
        cur = seen;
    }
}


struct PkArgs {
    const grace_b200_ray* rays; int n_packets;
    const float4* spheres; const int4* nodes; const int4* leaves; int n_nodes; const int* root_ptr;
    int* out_counts; float* out_cum; const int* offsets;
    int* hit_idx; float* hit_integral; float* hit_dist;
    int* unit_counter; int* err_flag; unsigned long long* prof;
};

// Pop the next stack entry that concerns this unit's rays (`subset`); entries whose mask
// has no ray of the unit are dropped here, so the walk never sees an empty mask.
#define PK_POP()                                                                         \
    do {                                                                                 \
        top = -1;                                                                        \
        while (sp > 0) {                                                                 \
            const int2 e_ = W.stack[--sp];                                               \
            if ((unsigned)e_.y & subset) { top = e_.x; top_mask = (unsigned)e_.y & subset; break; } \
        }                                                                                \
    } while (0)

// Slot word of a warp: 0 = nothing to steal here, s > 0 = a unit is walking, has run s - 1 steps and
// holds a stack entry, s < 0 = thief -s - 1 has asked for work.  Only the owner stores to the word
// (atomicExch at every check and when its unit ends, which also collects a pending request); a thief
// may only compare-and-swap a positive value it has just read into its request -- so a request can
// only land on a walking unit, which is certain to collect it.
//
// Victim side: thief `thief` asked.  Hand over the bottom of the stack if it holds a subtree worth a
// task.  Returns record << 6 | entries given, or -1 (nothing given; stack untouched).  On success
// the entries kept have moved down: W.stack[0, sp - given).
template <int MODE, int M4>
__device__ __noinline__ int pk_serve(PkWarp<MODE, M4>& W, const PkTasks& T, const int4* __restrict__ nodes, int n_nodes,
                                     int thief, int sp, int packet, unsigned subset, int prev_don, int lane)
{
    // The bottom entries up to and including the first one that spans enough leaves to be worth a task
    // (the tree is not balanced: the entry at the very bottom may be a single leaf with a large
    // subtree right above it).  They leave together -- taking an entry from the middle would put the
    // thief's terms in the middle of the victim's own.
    int rslot = -1, d = 0;
    const int look = min(sp, PK_DON_MAX);
    int2 e = make_int2(0, 0);
    int size = 0;
    if (lane < look) {
        e = W.stack[lane];
        size = 1;
        if (e.x < n_nodes) { const int4 n0 = __ldg(nodes + 4 * (size_t)e.x); size = n0.w - n0.z + 1; }
    }
    const unsigned big = __ballot_sync(0xffffffffu, lane < look && size >= (T.eager ? 1 : PK_DON_MINR));
    if (big) {
        d = __ffs(big);           // entries [0, d)
        if (lane == 0) { rslot = atomicAdd(T.lb + PK_LB_CREATED, 1); if (rslot >= T.records_cap) rslot = -1; }
        rslot = __shfl_sync(0xffffffffu, rslot, 0);
    }
    if (rslot >= 0) {
        int* rec = T.records + (size_t)rslot * PK_DREC_WORDS;
        if (lane == 0) {
            ((int4*)rec)[0] = make_int4(packet, (int)subset, d, 0);
            ((int4*)rec)[1] = make_int4(prev_don, -1, -1, 0);
        }
        if (lane < d) ((int2*)(rec + PK_DR_ENTRIES))[lane] = e;
        int2 ev[PK_STACK / 32 + 1];      // the entries kept move down by d
#pragma unroll
        for (int k = 0; k <= PK_STACK / 32; ++k) {
            const int i = lane + 32 * k;
            if (i + d < sp) ev[k] = W.stack[i + d];
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k <= PK_STACK / 32; ++k) {
            const int i = lane + 32 * k;
            if (i + d < sp) W.stack[i] = ev[k];
        }
        __threadfence();          // the record before the answer
    }
    __syncwarp();
    if (lane == 0) *(volatile int*)(T.resp + thief) = rslot >= 0 ? rslot : -2;
    return rslot >= 0 ? (rslot << 6) | d : -1;        // record and number of entries given (<= 32), packed
}

// Thief side (whole warp).  Returns a theft record to run as a task, or -1 when everything is done.
__device__ __noinline__ int pk_steal(const PkArgs& P, const PkTasks& T, int me, int n_units, int lane)
{
    unsigned rng = (unsigned)me * 2654435761u + (unsigned)clock();
#ifndef PK_RETRY0_V
#define PK_RETRY0_V 1024
#endif
    unsigned backoff = 128, retry = PK_RETRY0_V;
    const int min_age = max(T.budget, 1);
    PK_DBG(const unsigned long long w0 = pk_now(); if (lane == 0) atomicMin(&pk_dbg[4], w0);)
    for (int spins = 0;; ++spins) {
        // ---- scan: the longest-running unit among PK_SAMPLE runs of 32 consecutive slots (one 128-byte
        // line per load: thousands of thieves looking at random WORDS saturated the L2 and slowed the
        // walking units more than the stealing helped) ----
        int best = 0, bv = -1;
        const int n_lines = T.n_slots >> 5;
#pragma unroll
        for (int k = 0; k < PK_SAMPLE; ++k) {
            rng = rng * 1664525u + 1013904223u;
            const int v = (int)(((unsigned long long)rng * (unsigned)n_lines) >> 32) * 32 + lane;
            const int st = __ldcg(T.state + v);
            if (st > best) { best = st; bv = v; }
        }
        // ... and the units on the notice board: near the end of a launch the few units still walking
        // are the long-running ones, and a random sample of the slots mostly misses them
#pragma unroll
        for (int k = 0; k < PK_ADV / 32; ++k) {
            const int v = __ldcg(T.adv + k * 32 + lane) - 1;
            if (v >= 0) { const int st = __ldcg(T.state + v); if (st > best) { best = st; bv = v; } }
        }
        // every lane now holds the longest-running unit of its dozen candidates; one of the lanes that
        // found somebody old enough is drawn at random (if all thieves went for THE longest-running
        // unit they would collide on its slot word, and it answers one request per check)
        const unsigned cand = __ballot_sync(0xffffffffu, best > min_age);
        if (cand) {
            rng = rng * 1664525u + 1013904223u;
            int pick = (int)(((unsigned long long)rng * (unsigned)__popc(cand)) >> 32);
            unsigned cm = cand;
            while (pick-- > 0) cm &= cm - 1;
            bv = __shfl_sync(0xffffffffu, bv, __ffs(cm) - 1);
            int got = -3;         // -3: the request could not be posted
            if (lane == 0) {
                *(volatile int*)(T.resp + me) = -1;
                __threadfence();
                // the word changes at every check of its owner: a few attempts on its current value
                bool posted = false;
                for (int tries = 0; tries < 4 && !posted; ++tries) {
                    const int cur = *(volatile int*)(T.state + bv);
                    if (cur <= min_age) break;                 // gone, hidden, or somebody else's request is there
                    posted = atomicCAS(T.state + bv, cur, -(me + 1)) == cur;
                }
                if (posted) {
                    int polls = 0;
                    for (;;) {
                        got = *(volatile int*)(T.resp + me);
                        if (got != -1) break;
                        if (++polls > PK_SPIN_LIMIT) { atomicMax(P.err_flag, 3); got = -4; break; }
                        __nanosleep(500);
                    }
                }
            }
            got = __shfl_sync(0xffffffffu, got, 0);
            PK_DBG(if (lane == 0) atomicAdd(&pk_dbg[got >= 0 ? 16 : got == -2 ? 17 : 18], 1ull);)
            if (got >= 0) { __threadfence(); PK_DBG(if (lane == 0) atomicAdd(&pk_dbg[8], pk_now() - w0);) return got; }
            if (got == -4) return -1;
            // refused, or the word had moved on: look again, less and less eagerly (walking units answer one
            // request per check; thousands of thieves must not take their memory bandwidth meanwhile)
            __nanosleep(retry);
            if (retry < 32768) retry *= 2;
            continue;
        }
        // ---- nobody worth robbing in the sample: finished?  `finished` and `created` only grow and
        // finished <= n_units + created always holds, so equality with `created` read AFTER `finished`
        // proves that no unit is walking and none is on its way ----
        int done = 0;
        if (lane == 0) {
            const int f = *(volatile int*)(T.lb + PK_LB_FINISHED);
            __threadfence();
            const int c = *(volatile int*)(T.lb + PK_LB_CREATED);
            done = f >= n_units + min(c, T.records_cap);
            if (spins > PK_SPIN_LIMIT) { atomicMax(P.err_flag, 3); done = 1; }
        }
        if (__shfl_sync(0xffffffffu, done, 0)) { PK_DBG(if (lane == 0) atomicAdd(&pk_dbg[9], pk_now() - w0);) return -1; }
        PK_DBG(if (lane == 0) atomicAdd(&pk_dbg[19], 1ull);)
        __nanosleep(backoff);
        if (backoff < 4096) backoff *= 2;
    }
}

// Seven CTAs per SM (72 registers) for hit counts and column densities: hit counts carry no FIFO
// and no accumulators and gain 4 % (13.4 -> 12.8 ms); the column-density kernel still gains 1.3 %
// at 2^20 rays and 4.5 % at 2^17 despite a few spilled bytes.  Hit lists (more shared memory) stay at 6.
#ifndef PK_MIN_BLOCKS_LIGHT_V
#define PK_MIN_BLOCKS_LIGHT_V 7
#endif
// FOLD: the instantiation the fold launch runs (column densities only); the packet launch's own
// instantiation carries none of the fold code.
template <int MODE, int M4, bool PROF, bool FOLD>
__global__ void __launch_bounds__(PK_THREADS, FOLD ? 2 : ((MODE == MODE_COUNT || MODE == MODE_CUMULATIVE) && !PROF)
                                                  ? PK_MIN_BLOCKS_LIGHT_V : PK_MIN_BLOCKS)
trace_packet_kernel(const __grid_constant__ PkArgs P, const __grid_constant__ PkTasks T)
{
    using Warp = PkWarp<MODE, M4>;
    constexpr bool NEED_Q = Warp::NEED_Q;
    constexpr bool SUB = MODE != MODE_FILL;           // subtree donation; hit lists: ray-subset rounds
    constexpr bool CHAIN = MODE == MODE_CUMULATIVE || MODE == MODE_REC;   // units write to a pool: term chains + fold launch, or hit records
    constexpr bool REC = MODE == MODE_REC;            // hit lists in one traversal: EVERY unit records its hits
    extern __shared__ __align__(16) unsigned char pk_smem[];
    double2* s_table = (double2*)pk_smem;     // {T[i], T[i+1] - T[i]}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Warp& W = *((Warp*)(pk_smem + 52 * sizeof(double2)) + warp);
    if (NEED_Q) {
        for (int i = threadIdx.x; i < N_TABLE; i += PK_THREADS) {
            const double y0 = c_kernel_table[i], y1 = c_kernel_table[min(i + 1, N_TABLE - 1)];
            s_table[i] = make_double2(y0, __dsub_rn(y1, y0));
        }
        __syncthreads();
    }
    if (lane == 0) { W.pool = T.pool; W.pool_ctr = T.pool_ctr; W.pool_cap = T.pool_cap; W.ch_on = 0; W.ch_fail = 0; }
    const float4* __restrict__ spheres = P.spheres;
    const int4* __restrict__ nodes = P.nodes;
    const int4* __restrict__ leaves = P.leaves;
    const int n_nodes = P.n_nodes;
    const int root = __ldg(P.root_ptr);
    const unsigned lt = gb_lanemask_lt();
    // largest |coordinate| of any box plane of the tree (the root's two child boxes contain all others):
    // the planes are c -/+ h rounded to float, i.e. off by up to 2^-24 of their own magnitude, which the
    // padding of the slab test has to cover as well -- also for spheres much larger than the rays are long
    float tree_extent;
    {
        const int4* np = nodes + 4 * (size_t)root;
        const int4 n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
        tree_extent = fmaxf(fmaxf(fmaxf(fabsf(__int_as_float(n1.x)), fabsf(__int_as_float(n1.y))), fmaxf(fabsf(__int_as_float(n1.z)), fabsf(__int_as_float(n1.w)))),
                            fmaxf(fmaxf(fabsf(__int_as_float(n2.x)), fabsf(__int_as_float(n2.y))), fmaxf(fabsf(__int_as_float(n2.z)), fabsf(__int_as_float(n2.w)))));
        tree_extent = fmaxf(tree_extent, fmaxf(fmaxf(fabsf(__int_as_float(n3.x)), fabsf(__int_as_float(n3.y))), fmaxf(fabsf(__int_as_float(n3.z)), fabsf(__int_as_float(n3.w)))));
        if (!(tree_extent < 3.0e38f)) tree_extent = 0.0f;      // non-finite boxes: nothing sensible to add
    }
    const bool donating = SUB && !PROF && !FOLD && T.kind == PK_KIND_PACKETS && T.state != nullptr;
    const int n_units = FOLD ? __ldg(T.n_roots)
                      : T.kind == PK_KIND_TASKS ? min(__ldg(T.n_tasks_in), T.tasks_cap) : P.n_packets;

    PK_DBG(if (threadIdx.x == 0 && blockIdx.x == 0) atomicMin(&pk_dbg[7], pk_now());)
    const int me = blockIdx.x * PK_WARPS + warp;          // this warp's slot (state, mailbox, answer cell)

    for (;;) {
        int unit = -1;
        if (lane == 0) {
            // volatile pre-check keeps the ticket from running away
            if (*(volatile int*)P.unit_counter < n_units) unit = atomicAdd(P.unit_counter, 1);
            if (unit >= n_units) unit = -1;
        }
        unit = __shfl_sync(0xffffffffu, unit, 0);
        int my_task = -1;                 // >= 0: this unit is a stolen subtree (its theft record)
        if (unit < 0) {
            if (!donating) break;
            my_task = pk_steal(P, T, me, n_units, lane);      // every packet has been claimed: rob a walking unit
            if (my_task < 0) break;
        }
        int packet = unit;
        unsigned subset = 0xffffffffu;
        const int* rec = nullptr;
        if (my_task >= 0) {
            rec = T.records + (size_t)my_task * PK_DREC_WORDS;
            packet = __ldcg(rec + PK_DR_PACKET);
            subset = (unsigned)__ldcg(rec + PK_DR_SUBSET);
        } else if (T.kind == PK_KIND_TASKS) {
            const int2 t = T.tasks_in[unit];
            if (t.x < 0) continue;            // slot claimed but never filled
            rec = T.records + (size_t)t.x * PK_REC_WORDS;
            packet = rec[0];
            subset = (unsigned)t.y;
        } else if (FOLD) {
            packet = __ldg(&T.roots[unit].x);
        }
        const bool lane_on = (subset >> lane) & 1u;
        const int ray_index = packet * 32 + lane;
        const grace_b200_ray ray = P.rays[ray_index];
        __syncwarp();
        W.rays[2 * lane] = make_float4(ray.dx, ray.dy, ray.dz, ray.ox);
        W.rays[2 * lane + 1] = make_float4(ray.oy, ray.oz, ray.length, 0.f);
        PkSlab S;
        // 1/d, finite: with an infinite reciprocal fma(plane, 1/d, -(o +/- pad)/d) is inf - inf.  For
        // |d| < 2^-64 the ray is treated as parallel to the slab: t = +/-2^64 (plane - o -/+ pad),
        // whose sign -- all that matters then -- is decided well inside the padding.
        S.ix = pk_finite_rcp(ray.dx); S.iy = pk_finite_rcp(ray.dy); S.iz = pk_finite_rcp(ray.dz);
        S.len = ray.length;
        const float pad = 64.0f * 5.9604645e-8f *
                          (fabsf(ray.ox) + fabsf(ray.oy) + fabsf(ray.oz) + fabsf(ray.length) + tree_extent);
        // t_bottom = fma(b, inv, cb), t_top = fma(t, inv, ct) with the box grown by pad
        S.cbx = -(ray.ox + pad) * S.ix; S.ctx = -(ray.ox - pad) * S.ix;
        S.cby = -(ray.oy + pad) * S.iy; S.cty = -(ray.oy - pad) * S.iy;
        S.cbz = -(ray.oz + pad) * S.iz; S.ctz = -(ray.oz - pad) * S.iz;
        // the packet bound covers the unit's own rays only: lanes outside `subset` stand in
        // with a copy of one that is inside
        grace_b200_ray bray = ray;
        if (subset != 0xffffffffu) {
            const int src = __ffs(subset) - 1;
            const float b0 = __shfl_sync(0xffffffffu, ray.dx, src), b1 = __shfl_sync(0xffffffffu, ray.dy, src),
                        b2 = __shfl_sync(0xffffffffu, ray.dz, src), b3 = __shfl_sync(0xffffffffu, ray.ox, src),
                        b4 = __shfl_sync(0xffffffffu, ray.oy, src), b5 = __shfl_sync(0xffffffffu, ray.oz, src),
                        b6 = __shfl_sync(0xffffffffu, ray.length, src);
            if (!lane_on) { bray.dx = b0; bray.dy = b1; bray.dz = b2; bray.ox = b3; bray.oy = b4; bray.oz = b5; bray.length = b6; }
        }
        const PacketBound B = packet_bound(bray);
        // all 32 rays share one origin?  (comparisons AFTER the shuffles in packet_bound:
        // every lane must execute every shuffle)
        const bool common = __all_sync(0xffffffffu, (ray.ox == B.ocx) & (ray.oy == B.ocy) & (ray.oz == B.ocz));
        PkAcc A;
        A.count = 0; A.cum = 0.0f; A.cursor = 0; A.qn = 0; A.room = PK_QD;
        A.hit_idx = P.hit_idx; A.hit_integral = P.hit_integral; A.hit_dist = P.hit_dist;
        int sp = 0;
        int top = root;                  // top of stack in a register; -1 = empty, <= -2 = FOLD entry
        unsigned top_mask = 0xffffffffu;
        if (CHAIN && lane == 0) {
            W.ch_on = 0; W.ch_fail = 0; W.ch_cur = -1; W.ch_head = -1;
            if (REC) { W.ch_off = 0; W.ch_nb = 0; W.ch_unit = my_task >= 0 ? -1 - my_task : packet; }
        }
        if (SUB && my_task >= 0) {
            // stolen stack entries, for all rays of the unit they were taken from
            sp = __ldcg(rec + PK_DR_N);
            if (lane < sp) W.stack[lane] = __ldcg((const int2*)(rec + PK_DR_ENTRIES) + lane);
            if (CHAIN && !REC && lane == 0) W.ch_on = 1;
            __syncwarp();
            PK_POP();
        } else if (FOLD) {
            // a robbed packet: its own sum, then what was stolen from it, latest theft first.  A fold
            // cursor is the stack entry {-2 - theft record, all lanes}
            A.cum = __ldg(T.root_cum + (size_t)unit * 32 + lane);
            top = -2 - __ldg(&T.roots[unit].y);
        } else if (rec) {                // hit lists: the whole traversal state, for the rays in `subset`
            sp = __ldcg(rec + 2); top = __ldcg(rec + 3); top_mask = (unsigned)__ldcg(rec + 4) & subset;
            for (int i = lane; i < sp; i += 32) W.stack[i] = __ldcg((const int2*)(rec + 8) + i);
            __syncwarp();
            if (top_mask == 0u) PK_POP();
            A.cum = __int_as_float(__ldcg(rec + 8 + 2 * PK_STACK + lane));
            A.count = __ldcg(rec + 8 + 2 * PK_STACK + 32 + lane);
            A.cursor = __ldcg(rec + 8 + 2 * PK_STACK + 64 + lane);
        } else if (MODE == MODE_FILL) {
            A.cursor = P.offsets[ray_index];
        }
        __syncwarp();

        unsigned long long pf_nodes = 0, pf_leaves = 0, pf_staged = 0, pf_kept = 0, pf_l1 = 0, pf_l2 = 0, pf_l3 = 0;
        (void)pf_l1; (void)pf_l2; (void)pf_l3;
        const long long pf_t0 = PROF ? clock64() : 0;
        const int guard0 = 2 * n_nodes + 8;   // a depth-first walk enters each node at most once
        int guard = guard0;
        bool suspended = false, aborted = false;
        int don_head = -1;                // what has been stolen from this unit (latest theft first)
        int next_check = T.budget, hold_until = 0;
        PK_DBG(unsigned long long dbg_recs = 0; const unsigned long long dbg_t0 = pk_now();)
        for (;;) {
            if (MODE == MODE_CUMULATIVE && W.ch_fail) { aborted = true; break; }      // the chunk pool ran dry
            // ---- every PK_CHECK_EVERY steps: publish progress, collect and answer a thief's request ----
            if (donating && guard0 - guard >= next_check) {
                const int steps = guard0 - guard;
                next_check = steps + PK_CHECK_EVERY;
                int old = 0;
                if (lane == 0) {
                    old = atomicExch(T.state + me, (sp > 0 && steps >= hold_until) ? steps + 1 : 0);
                    if (steps >= PK_ADV_AGE) T.adv[me & (PK_ADV - 1)] = me + 1;
                }
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old < 0) {
                    const int gave = pk_serve<MODE, M4>(W, T, nodes, n_nodes, -old - 1, sp, packet, subset, don_head, lane);
                    const int rslot = gave >= 0 ? gave >> 6 : -1, given = gave & 63;
                    PK_DBG(if (lane == 0) atomicAdd(&pk_dbg[rslot >= 0 ? 10 : 11], 1ull);)
                    if (rslot >= 0) { don_head = rslot; sp -= given; next_check = steps + 2; }       // in demand: look again soon
                    else hold_until = steps + 8 * PK_CHECK_EVERY;       // nothing worth a task down there: stay out of sight a while
                }
            }
            // ---- hit lists: suspend an over-budget traversal, resume it as ray-subset tasks ----
            // Splitting costs SIMD width (a task owns fewer rays), so it only pays when warps would
            // otherwise idle: a unit is suspended once it has run `budget` steps AND the ticket counter
            // shows that no unclaimed unit is left AND it is heavier than most.
            if (!SUB && T.tasks_out != nullptr && guard0 - guard >= T.budget && top >= 0) {
                bool split_now = true;
                if (!T.eager) {
                    split_now = *(volatile int*)P.unit_counter >= n_units;
                    if (split_now && T.n_done) {
                        const long long nd = *(volatile int*)T.n_done;
                        const unsigned long long ss = *(volatile unsigned long long*)T.sum_steps;
                        if (nd > 0) split_now = 4ll * (guard0 - guard) * nd >= (long long)PK_AVG_FACTOR_X4 * (long long)ss;
                        else split_now = guard0 - guard >= 8 * T.budget;      // nothing to compare with yet
                    }
                }
                if (split_now) {
                    // rays per child task 8, 2, 1 by round (one launch per round)
                    const int child_width = T.child_width;
                    const unsigned bmask0 = (1u << child_width) - 1u;
                    unsigned blocks = 0;          // bit b: some ray of lane block b belongs to this unit
                    for (int b = 0; b * child_width < 32; ++b)
                        if (subset & (bmask0 << (b * child_width))) blocks |= 1u << b;
                    const int nchild = __popc(blocks);
                    int rslot = -1, tslot = -1;
                    if (nchild > 1 && lane == 0) {
                        tslot = pk_reserve(T.n_tasks_out, nchild, T.tasks_cap);
                        if (tslot >= 0) {
                            rslot = pk_reserve(T.n_records, 1, T.records_cap);
                            if (rslot < 0)         // no record: publish nothing in the claimed task slots
                                for (int c = 0; c < nchild; ++c) T.tasks_out[tslot + c] = make_int2(-1, 0);
                        }
                    }
                    rslot = __shfl_sync(0xffffffffu, rslot, 0);
                    tslot = __shfl_sync(0xffffffffu, tslot, 0);
                    if (rslot >= 0) {
                        if (NEED_Q) pk_flush<MODE, M4>(W, A, lane, lt, s_table);
                        int* r = T.records + (size_t)rslot * PK_REC_WORDS;
                        if (lane == 0) { r[0] = packet; r[1] = (int)subset; r[2] = sp; r[3] = top; r[4] = (int)top_mask; r[5] = 0; }
                        for (int i = lane; i < sp; i += 32) ((int2*)(r + 8))[i] = W.stack[i];
                        r[8 + 2 * PK_STACK + lane] = __float_as_int(A.cum);
                        r[8 + 2 * PK_STACK + 32 + lane] = A.count;
                        r[8 + 2 * PK_STACK + 64 + lane] = A.cursor;
                        if (lane < nchild) {      // lane c publishes the c-th non-empty block
                            unsigned bb = blocks;
                            for (int c = 0; c < lane; ++c) bb &= bb - 1;
                            const int b = __ffs(bb) - 1;
                            T.tasks_out[tslot + lane] = make_int2(rslot, (int)(subset & (bmask0 << (b * child_width))));
                        }
                        suspended = true;
                        break;
                    }
                }
            }
            // ---- fold unit: the next pending item is a theft record ----
            if (FOLD && top <= -2) {
                PK_DBG(++dbg_recs;)
                pk_flush<MODE, M4>(W, A, lane, lt, s_table);      // hits of subtrees walked in place come first
                const int4* drec = (const int4*)(T.records + (size_t)(-2 - top) * PK_DREC_WORDS);
                const int4 d0 = __ldcg(drec), d1 = __ldcg(drec + 1);      // packet, subset, entries, - | older, head, thefts, status
                if (sp + 1 + PK_DON_MAX > PK_STACK) { if (lane == 0) atomicMax(P.err_flag, 1); top = -1; sp = 0; continue; }
                // after this task and everything stolen from it: the previous theft from the same unit
                if (d1.x >= 0) {
                    if (lane == 0) { W.stack[sp] = make_int2(-2 - d1.x, -1); asm volatile("prefetch.global.L2 [%0];" :: "l"(T.records + (size_t)d1.x * PK_DREC_WORDS)); }
                    ++sp;
                }
                if (d1.w == 1) {
                    // what was stolen from the task comes right after its own chain: have that record on its way
                    if (d1.z >= 0 && lane == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(T.records + (size_t)d1.z * PK_DREC_WORDS));
                    A.cum = pk_fold_chain(T.pool, d1.y, A.cum, lane, lt);
                    if (d1.z >= 0) { if (lane == 0) W.stack[sp] = make_int2(-2 - d1.z, -1); ++sp; }
                } else {                      // aborted or never run: walk its subtree here, in place (what was
                                              // stolen from it before it aborted is covered by that walk and ignored)
                    if (lane < d0.z) W.stack[sp + lane] = __ldcg((const int2*)((const int*)drec + PK_DR_ENTRIES) + lane);
                    sp += d0.z;
                }
                __syncwarp();
                PK_POP();
                continue;
            }
            // ---- phase A: find the next leaves (at most PK_BATCH, in ascending order) ----
            int nb = 0;
            int b_leaf = 0;               // lane b holds batch entry b
            unsigned b_mask = 0;
            while (top >= 0 && nb < PK_BATCH) {
                if (--guard < 0) { if (lane == 0) atomicMax(P.err_flag, 2); top = -1; sp = 0; break; }
                if (top < n_nodes) {
                    if (PROF) ++pf_nodes;
                    const int4* np = nodes + 4 * (size_t)top;
                    const int4 n0 = __ldg(np + 0);
                    const int4 n1 = __ldg(np + 1);
                    const int4 n2 = __ldg(np + 2);
                    const int4 n3 = __ldg(np + 3);
                    const bool hitL = pk_slab(S, n1.x, n1.y, n1.z, n1.w, n3.x, n3.y);
                    const bool hitR = pk_slab(S, n2.x, n2.y, n2.z, n2.w, n3.z, n3.w);
                    // a lane that missed an ancestor's box cannot hit anything below it
                    const unsigned mL = __ballot_sync(0xffffffffu, hitL) & top_mask;
                    const unsigned mR = __ballot_sync(0xffffffffu, hitR) & top_mask;
                    if (mL && mR) {
                        if (sp >= PK_STACK) { if (lane == 0) atomicMax(P.err_flag, 1); top = -1; sp = 0; continue; }
                        W.stack[sp++] = make_int2(n0.y, (int)mR);
                        top = n0.x; top_mask = mL;
                    } else if (mL) {
                        top = n0.x; top_mask = mL;
                    } else if (mR) {
                        top = n0.y; top_mask = mR;
                    } else {
                        PK_POP();
                    }
                } else {
                    if (lane == nb) { b_leaf = top - n_nodes; b_mask = top_mask; }
                    ++nb;
                    PK_POP();
                }
            }
            if (nb == 0) {
                if (FOLD && top <= -2) continue;
                break;
            }
            // ---- phase B: all leaf records in one round trip, then all spheres in one ----
            int2 lf = make_int2(0, 0);
            if (lane < nb && b_mask) lf = __ldg((const int2*)(leaves + b_leaf));
            for (int slot = 0; slot < nb; ++slot) {
                const int first = __shfl_sync(0xffffffffu, lf.x, slot);
                const int cnt = min(__shfl_sync(0xffffffffu, lf.y, slot), M4);
                for (int i = lane; i < cnt; i += 32)
                    cp_async16(&W.raw[slot * M4 + i], spheres + first + i);
            }
            cp_async_wait_all();
            __syncwarp();
            // ---- phase C: cull, compact and test leaf by leaf (ascending order) ----
            for (int slot = 0; slot < nb; ++slot) {
                const int first = __shfl_sync(0xffffffffu, lf.x, slot);
                const int cnt = min(__shfl_sync(0xffffffffu, lf.y, slot), M4);
                const unsigned leaf_mask = __shfl_sync(0xffffffffu, b_mask, slot);
                if (leaf_mask == 0u) continue;
                int n_kept = 0;
                for (int base = 0; base < cnt; base += 32) {
                    const int i = base + lane;
                    bool keep = false;
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < cnt) {
                        s = W.raw[slot * M4 + i];
                        keep = !B.enabled || packet_may_hit(B, s);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const int dst = n_kept + __popc(m & lt);
                        if (NEED_Q) W.ir[dst] = __fdiv_rn(1.0f, s.w);
                        if (NEED_I_(MODE)) W.idx[dst] = first + i;
                        s.w = __fmul_rn(s.w, s.w);
                        if (common) {
                            s.x = __fsub_rn(s.x, ray.ox); s.y = __fsub_rn(s.y, ray.oy); s.z = __fsub_rn(s.z, ray.oz);
                        }
                        W.prims[dst] = s;
                    }
                    n_kept += __popc(m);
                }
                __syncwarp();
                if (PROF) { ++pf_leaves; pf_staged += cnt; pf_kept += n_kept; }
                const int k_active = __popc(leaf_mask);
#ifdef PK_PROF_LANES      // diagnostic build: lane efficiency of the leaf tests instead of the usual per-packet record
                if (PROF) {
                    if (PK_SPARSE_NUM * k_active < PK_SPARSE_DEN * n_kept) pf_l3 += (unsigned long long)n_kept * k_active;
                    else { pf_l1 += (unsigned long long)n_kept * k_active; pf_l2 += n_kept; }
                }
#endif
                if (PK_SPARSE_NUM * k_active < PK_SPARSE_DEN * n_kept) {
                    if (common) pk_leaf_sparse<MODE, M4, true>(W, leaf_mask, n_kept, lane, lt, A, s_table);
                    else pk_leaf_sparse<MODE, M4, false>(W, leaf_mask, n_kept, lane, lt, A, s_table);
                } else {
                    if (common) pk_leaf_dense<MODE, M4, true>(W, n_kept, ray, lane, lt, lane_on, A, s_table);
                    else pk_leaf_dense<MODE, M4, false>(W, n_kept, ray, lane, lt, lane_on, A, s_table);
                }
                __syncwarp();
            }
        }
        if (suspended) continue;
        PK_DBG(if (FOLD && lane == 0) { atomicAdd(&pk_dbg[22], dbg_recs); atomicMax(&pk_dbg[23], dbg_recs); atomicMax(&pk_dbg[24], pk_now() - dbg_t0); atomicAdd(&pk_dbg[25], pk_now() - dbg_t0); atomicAdd(&pk_dbg[26], (unsigned long long)(guard0 - guard)); })
        if (NEED_Q && !aborted) pk_flush<MODE, M4>(W, A, lane, lt, s_table);
        if (donating) {                      // no longer walking: thieves look elsewhere, a late request gets "nothing"
            if (lane == 0) {
                const int old = atomicExch(T.state + me, 0);
                if (old < 0) *(volatile int*)(T.resp + (-old - 1)) = -2;
                if (guard0 - guard >= PK_ADV_AGE && T.adv[me & (PK_ADV - 1)] == me + 1) T.adv[me & (PK_ADV - 1)] = 0;
            }
        }
        if (REC && !FOLD && W.ch_fail && lane == 0) atomicExch(T.lb + PK_LB_OVERFLOW, 1);     // pool dry: counts are right, lists incomplete
        if (SUB && my_task >= 0) {           // a stolen subtree
            if (REC) {                       // hit records: its own hits per ray (the stack entries are no longer needed) and its thefts
                const int start = pk_rec_close<MODE, M4>(W, T.n_roots, lane);
                int* r = T.records + (size_t)my_task * PK_DREC_WORDS;
                r[PK_DR_ENTRIES + lane] = A.count;
                if (lane == 0) { r[PK_DR_HEAD] = start; r[PK_DR_DONS] = don_head; r[PK_DR_STATUS] = 1; }
            } else if (CHAIN) {              // publish its chain (or its failure) and what was stolen from it
                if (W.ch_fail) aborted = true;
                pk_chain_close<MODE, M4>(W, lane);
                if (lane == 0) {
                    int* r = T.records + (size_t)my_task * PK_DREC_WORDS;
                    r[PK_DR_HEAD] = W.ch_head; r[PK_DR_DONS] = don_head; r[PK_DR_STATUS] = aborted ? 2 : 1;
                }
            }
            if ((MODE == MODE_COUNT || REC) && lane_on && A.count) atomicAdd(P.out_counts + ray_index, A.count);
        } else if (SUB && don_head >= 0) {   // a robbed packet
            if ((MODE == MODE_COUNT || REC) && lane_on) atomicAdd(P.out_counts + ray_index, A.count);   // its tasks add to the same (zeroed) cell
            if (REC) {                       // its own hits per ray and the thefts, for rec_resolve_kernel
                const int start = pk_rec_close<MODE, M4>(W, T.n_roots, lane);
                ((int*)T.root_cum)[ray_index] = A.count;
                if (lane == 0) T.roots[packet] = make_int2(start, don_head);
            } else if (CHAIN) {              // its own sum goes to a root slot; the fold launch adds the rest
                int slot = 0;
                if (lane == 0) slot = atomicAdd(T.n_roots, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                T.root_cum[(size_t)slot * 32 + lane] = A.cum;
                if (lane == 0) T.roots[slot] = make_int2(packet, don_head);
            }
        } else {
            if ((MODE == MODE_COUNT || REC) && lane_on) P.out_counts[ray_index] = A.count;
            if (MODE == MODE_CUMULATIVE && lane_on) P.out_cum[ray_index] = A.cum;
            if (REC) {
                const int start = pk_rec_close<MODE, M4>(W, T.n_roots, lane);
                if (lane == 0) T.roots[packet] = make_int2(start, -1);
            }
        }
        if (donating) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(T.lb + PK_LB_FINISHED, 1);
            PK_DBG(if (lane == 0) {
                const unsigned long long st = (unsigned long long)(guard0 - guard);
                if (my_task >= 0) { atomicAdd(&pk_dbg[0], st); atomicMax(&pk_dbg[1], st); atomicMax(&pk_dbg[6], pk_now()); if (st < 16) atomicAdd(&pk_dbg[14], 1ull); }
                else { atomicAdd(&pk_dbg[2], st); atomicMax(&pk_dbg[3], st); atomicMax(&pk_dbg[5], pk_now()); }
            })
        }
        if (!SUB && T.n_done && lane == 0) { atomicAdd(T.sum_steps, (unsigned long long)(guard0 - guard)); atomicAdd(T.n_done, 1); }
        if (PROF && lane == 0) {
            atomicAdd(P.prof + 0, pf_nodes); atomicAdd(P.prof + 1, pf_leaves);
            atomicAdd(P.prof + 2, pf_staged); atomicAdd(P.prof + 3, pf_kept);
            unsigned long long* pp = P.prof + 4 + 4 * (size_t)packet;   // per-packet record
            pp[0] = (unsigned long long)(clock64() - pf_t0); pp[1] = pf_nodes; pp[2] = pf_leaves; pp[3] = pf_kept;
#ifdef PK_PROF_LANES
            pp[1] = pf_l1; pp[2] = pf_l2; pp[3] = pf_l3;
#endif
        }
    }
}

template <int MODE, int M4>
constexpr size_t packet_smem_bytes()
{
    return 52 * sizeof(double2) + PK_WARPS * sizeof(PkWarp<MODE, M4>);
}
