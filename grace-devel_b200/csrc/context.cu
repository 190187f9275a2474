// context.cu -- context, workspace arena and error plumbing of the C ABI.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

static thread_local char g_err[512] = "";

int gb_set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int gb_cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    return gb_set_error(GRACE_B200_ECUDA, "CUDA error %d (%s) in %s at %s:%d",
                        (int)e, cudaGetErrorString(e), what, file, line);
}

void* gb_workspace(grace_b200_ctx* ctx, size_t bytes)
{
    if (bytes > GB_WS_HEAD) ++ctx->ws_epoch;       // whatever a trace call recorded beyond the head is gone
    if (bytes <= ctx->ws_bytes) return ctx->ws;
    ++ctx->ws_epoch;
    size_t want = gb_align(bytes + bytes / 8, 1 << 20);
    // Growth: drain outstanding work that may still use the old arena.
    if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
    if (ctx->ws) cudaFree(ctx->ws);
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        gb_set_error(GRACE_B200_ENOMEM, "workspace cudaMalloc(%zu) failed: %s", want,
                     cudaGetErrorString(e));
        return nullptr;
    }
    ctx->ws = (char*)p;
    ctx->ws_bytes = want;
    return p;
}

extern "C" {

const char* grace_b200_last_error(void) { return g_err; }
const char* grace_b200_version(void) { return "grace_b200 0.1 (sm_100a)"; }

int grace_b200_create(grace_b200_ctx** out, int device)
{
    GB_REQUIRE(out != nullptr, GRACE_B200_EINVAL, "ctx out pointer is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return gb_set_error(GRACE_B200_ECUDA,
                            "no CUDA device available (%s); grace_b200 has no CPU fallback",
                            cudaGetErrorString(e));
    GB_REQUIRE(device >= 0 && device < count, GRACE_B200_EINVAL, "device %d out of range", device);
    GB_CUDA(cudaSetDevice(device));
    grace_b200_ctx* ctx = new (std::nothrow) grace_b200_ctx();
    GB_REQUIRE(ctx != nullptr, GRACE_B200_ENOMEM, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    // Part of the (126 MB on B200) L2 can be set aside for data marked "persisting": the trace launches mark
    // the node array (54 MB at 2^24 particles), whose records every step of every packet waits for.
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    {
        const char* op = getenv("GRACE_B200_ONE_PASS_LISTS");
        if (op) ctx->one_pass_lists = atoi(op);
        const char* e = getenv("GRACE_B200_L2_PERSIST");
        ctx->l2_persist = e ? atoi(e) : 0;
        if (ctx->l2_persist && ctx->l2_persist_max)
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max);
    }
    GB_CUDA(cudaMalloc((void**)&ctx->d_scalars, GB_SC_COUNT * sizeof(int)));
    GB_CUDA(cudaMemset(ctx->d_scalars, 0, GB_SC_COUNT * sizeof(int)));
    GB_CUDA(cudaMallocHost((void**)&ctx->h_pinned, GB_SC_COUNT * sizeof(int)));
    memset(ctx->h_pinned, 0, GB_SC_COUNT * sizeof(int));
    *out = ctx;
    return GRACE_B200_OK;
}

int grace_b200_destroy(grace_b200_ctx* ctx)
{
    if (!ctx) return GRACE_B200_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux) grace_b200_destroy(ctx->aux);
    if (ctx->tile_mem) cudaFree(ctx->tile_mem);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage_host[i]) cudaFreeHost(ctx->stage_host[i]);
        if (ctx->stage_done[i]) cudaEventDestroy(ctx->stage_done[i]);
    }
    delete ctx;
    return GRACE_B200_OK;
}

int grace_b200_reserve(grace_b200_ctx* ctx, size_t bytes)
{
    GB_REQUIRE(ctx != nullptr, GRACE_B200_EINVAL, "ctx is NULL");
    return gb_workspace(ctx, bytes) ? GRACE_B200_OK : GRACE_B200_ENOMEM;
}

size_t grace_b200_workspace_bytes(const grace_b200_ctx* ctx) { return ctx ? ctx->ws_bytes : 0; }

} // extern "C"
