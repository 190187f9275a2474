// common.cuh -- shared host/device helpers for the grace_b200 CUDA sources.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "grace_b200.h"

// ---------------------------------------------------------------------------
// Context: device, SM count, and a grow-only device workspace arena.
// ---------------------------------------------------------------------------
struct grace_b200_ctx {
    int device = 0;
    int sm_count = 148;
    char* ws = nullptr;        // workspace arena
    size_t ws_bytes = 0;
    int* d_scalars = nullptr;  // small persistent device scalars (tickets, counts)
    int* h_pinned = nullptr;   // pinned host mirror for count read-backs
    int last_n_leaves_valid = 0;
    // streamed hit lists (hits.cu): helper context + stream for the sort side, reusable tile buffers
    grace_b200_ctx* aux = nullptr;
    cudaStream_t aux_stream = nullptr;
    char* tile_mem = nullptr;
    size_t tile_bytes = 0;
    // one-pass hit lists (trace.cu, hits.cu): the hits the count call recorded in the workspace, for the fill call
    unsigned long long ws_epoch = 0;        // bumped whenever a call takes more of the arena than its small head
    int rec_valid = 0;
    unsigned long long rec_epoch = 0;
    const void* rec_rays = nullptr; size_t rec_n_rays = 0; const void* rec_offsets = nullptr;
    alignas(16) unsigned char rec_blob[512] = {};   // launch arguments of the copy launch
    size_t rec_units = 0;                   // packets of the last recording
    size_t rec_pool_learned = 0;            // what a recording that overflowed its pool would have needed
    size_t rec_pool_hint = 0;               // bytes of hit pool the next recording should use (0 = from the ray count)
    int one_pass_lists = 1;
    size_t leaves_stage_n = 0;       // > 0: the workspace still holds the leaf-level deltas of an albvh_leaves call
    int leaves_stage_delta_type = 0;
    int trace_mode = GRACE_B200_TRACE_PACKET;
    int trace_budget = 64;     // traversal steps before a unit may donate / be suspended (0 = never)
    int trace_budget_set = 0;  // the caller has chosen it (else 16 for work stealing, 64 for the hit-list rounds)
    // L2 residency for the node array during traversal (the dependent node fetch is the latency chain)
    size_t l2_persist_max = 0, l2_window_max = 0;
    int l2_persist = 0;             // 1: set an access-policy window on the nodes around trace launches
    size_t trace_pool_bytes = 0;    // term pool of the column-density load balancing (0 = sized from the ray count)
    // file-loader staging (gadget_io.cu): two pinned host buffers + completion events, kept
    char* stage_host[2] = { nullptr, nullptr };
    size_t stage_bytes = 0;
    cudaEvent_t stage_done[2] = { nullptr, nullptr };
};

// d_scalars layout (ints)
enum {
    GB_SC_TICKET0 = 0,      // generic block tickets (self-resetting)
    GB_SC_TICKET1 = 1,
    GB_SC_NLEAVES = 2,      // leaf count of the last build
    GB_SC_TICKET2 = 3,      // group tickets of the node build
    GB_SC_TRACE_CTR = 4,    // packet scheduler counter
    GB_SC_ERRFLAG = 5,      // device-side error flag (trace stack overflow)
    GB_SC_TICKET3 = 6,      // segment tickets of the segmented scans (zeroed before every launch; TICKET0 must be 0 between launches)
    GB_SC_TOTAL64 = 8,      // 64-bit total (2 ints), 8-byte aligned
    GB_SC_TASKS = 32,       // trace load balancing: record count + two task-list counts
    GB_SC_CLASS = 40,       // segmented sort class counters (16 ints) + XL total (2 ints, 8-byte aligned)
    GB_SC_LB = 64,          // trace work donation: {finished, tail, head, idle} (one 16-byte load), then
                            // [4] records, [5] chunks taken, [6] root records
    GB_SC_MINMAX = 80,      // min/max of four components (8 floats) for the host-returning bounds call
    GB_SC_COUNT = 96
};

int gb_set_error(int code, const char* fmt, ...);
int gb_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
// Returns a pointer to at least `bytes` of workspace (256-byte aligned), growing
// the arena if needed (growth synchronises the device).  nullptr on failure.
void* gb_workspace(grace_b200_ctx* ctx, size_t bytes);
// The first GB_WS_HEAD bytes of the arena are for small helpers (scan states ...): a request that fits
// there leaves whatever a trace call recorded behind it intact.
constexpr size_t GB_WS_HEAD = 65536;
// one-pass hit lists (trace.cu): record the hits while counting / copy them out at the offsets
int gb_trace_record_f4(grace_b200_ctx* ctx, const grace_b200_ray* d_rays, size_t n_rays, const float* d_spheres4, size_t n,
                       const grace_b200_tree* tree, int* d_counts, cudaStream_t st);
int gb_trace_resolve_recorded(grace_b200_ctx* ctx, const int* d_offsets, cudaStream_t st);
int gb_trace_copy_recorded(grace_b200_ctx* ctx, const int* d_offsets, int* d_idx, float* d_integ, float* d_dist, cudaStream_t st);

#define GB_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t gb_e_ = (call);                                           \
        if (gb_e_ != cudaSuccess) return gb_cuda_fail(gb_e_, #call, __FILE__, __LINE__); \
    } while (0)

#define GB_LAUNCH_CHECK() GB_CUDA(cudaPeekAtLastError())

#define GB_REQUIRE(cond, code, ...)                                           \
    do { if (!(cond)) return gb_set_error((code), __VA_ARGS__); } while (0)

static inline size_t gb_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Simple bump allocator over one gb_workspace() block.
struct GbArena {
    char* base; size_t off, cap;
    GbArena(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T> T* take(size_t count) {
        size_t bytes = gb_align(count * sizeof(T));
        T* p = (T*)(base + off);
        off += bytes;
        return p;
    }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned gb_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned gb_lanemask_lt() {
    unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// Streaming (read-once) 128-bit load / store: keep L1 for data that is reused.
__device__ __forceinline__ float4 gb_ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned gb_ld_volatile_u32(const unsigned* p) {
    unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ unsigned long long gb_ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void gb_st_volatile_u32(unsigned* p, unsigned v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void gb_st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// L2-coherent (L1-bypassing) 128-bit load for data written by other SMs in the same launch.
__device__ __forceinline__ int4 gb_ld_cg_i4(const int4* p) {
    int4 r;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
#endif
