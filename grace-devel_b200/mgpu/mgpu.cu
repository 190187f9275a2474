// mgpu.cu -- one host process, every GPU of the node: replicated tree, rays dealt in 32-aligned
// tiles, NCCL broadcast of the inputs and gather of the per-ray outputs (include/grace_b200_mgpu.h).
//
// The reference has no counterpart (its profilers take a device id,
// tests/profile_one_to_many_rays_gadget/profile_one_to_many_rays_gadget.cu:43-52); what is
// reproduced on every device is its single-GPU path through the grace_b200 C ABI.  One worker
// thread per device issues that device's calls, so the per-device builds and traces run
// concurrently although several of those calls synchronise their stream.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "grace_b200_mgpu.h"

// this library's own error text (libgrace_b200's internals are not exported)
static thread_local char g_mg_err[512] = "";
static std::string g_mg_err_shared;          // last failure of a worker thread, for the calling thread
static int gb_set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_mg_err, sizeof(g_mg_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

constexpr size_t TILE = 4096;      // rays per tile; a multiple of the packet width (32)

struct Dev {
    int id = 0;
    grace_b200_ctx* ctx = nullptr;
    cudaStream_t st = nullptr;
    ncclComm_t comm = nullptr;
    float* spheres = nullptr;            // n x float4, sorted
    int4* nodes = nullptr; size_t nodes_cap = 0;   // int4 units
    int4* leaves = nullptr; size_t leaves_cap = 0;
    int* root = nullptr;
    grace_b200_ray* rays_full = nullptr; size_t rays_full_cap = 0;   // the whole ray set (devices other than 0)
    grace_b200_ray* rays = nullptr; size_t rays_cap = 0;     // this device's tiles, concatenated
    float* out = nullptr; size_t out_cap = 0;                // 4 bytes per local ray
    size_t n_local = 0;
    int rc = 0;
    std::string err;
};

} // namespace

struct grace_b200_mgpu {
    std::vector<Dev> d;
    size_t n = 0;               // particles
    int n_leaves = 0, max_per_leaf = 0;
    // device 0: whole ray set / gathered results in rank-major order / results in ray order
    grace_b200_ray* rays_all = nullptr; size_t rays_all_cap = 0;
    float* gathered = nullptr; float* ordered = nullptr; size_t res_cap = 0;
};

namespace {

#define MG_CUDA(call)                                                                                      \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return gb_set_error(GRACE_B200_ECUDA, "CUDA error %d (%s) in %s at %s:%d", \
                                                                                (int)e_, cudaGetErrorString(e_), #call, __FILE__, __LINE__); } while (0)
#define MG_NCCL(call)                                                                                      \
    do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return gb_set_error(GRACE_B200_ECUDA, "NCCL error %d (%s) in %s at %s:%d", \
                                                                                (int)r_, ncclGetErrorString(r_), #call, __FILE__, __LINE__); } while (0)

// Run f(device) on one thread per device; the first non-zero return code wins.
int for_each_device(grace_b200_mgpu* mg, const std::function<int(Dev&)>& f)
{
    std::vector<std::thread> th;
    for (auto& dv : mg->d)
        th.emplace_back([&dv, &f]() {
            cudaSetDevice(dv.id);
            g_mg_err[0] = 0;
            dv.rc = f(dv);
            if (dv.rc) dv.err = g_mg_err[0] ? g_mg_err : grace_b200_last_error();     // error texts are per thread
        });
    for (auto& t : th) t.join();
    for (auto& dv : mg->d) if (dv.rc) return gb_set_error(dv.rc, "device %d: %s", dv.id, dv.err.c_str());
    return GRACE_B200_OK;
}

template <typename T>
int ensure(T** p, size_t* cap, size_t need)
{
    if (*cap >= need && *p) return GRACE_B200_OK;
    if (*p) MG_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    MG_CUDA(cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return GRACE_B200_OK;
}

size_t n_tiles_of(size_t n_rays) { return (n_rays + TILE - 1) / TILE; }
// rays device `rank` of `world` owns: its tiles t = rank, rank + world, ...
size_t local_count(size_t n_rays, int rank, int world)
{
    size_t c = 0;
    for (size_t t = rank; t < n_tiles_of(n_rays); t += world) c += (t + 1) * TILE <= n_rays ? TILE : n_rays - t * TILE;
    return c;
}

// dst (this device's tiles, concatenated) <- src (all rays), 28-byte records moved as 7 floats
__global__ void take_tiles_kernel(const float* __restrict__ all, float* __restrict__ local, size_t n_rays, int rank, int world)
{
    for (size_t t = rank, lt = 0; t * TILE < n_rays; t += world, ++lt) {
        const size_t m = (t + 1) * TILE <= n_rays ? TILE : n_rays - t * TILE;
        const float* s = all + 7 * t * TILE;
        float* d = local + 7 * lt * TILE;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 7 * m; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
    }
}

// ordered[ray] <- gathered[rank-major]; every rank's block is padded to `per` results
__global__ void reorder_kernel(const float* __restrict__ gathered, float* __restrict__ ordered, size_t n_rays, int world, size_t per)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_rays; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i / TILE, o = i % TILE;
        const int rank = (int)(t % world);
        const size_t lt = t / world;
        ordered[i] = gathered[rank * per + lt * TILE + o];
    }
}

struct Ev {
    cudaEvent_t e[5];
    Ev() { for (auto& x : e) cudaEventCreate(&x); }
    ~Ev() { for (auto& x : e) cudaEventDestroy(x); }
};

int trace_any(grace_b200_mgpu* mg, const grace_b200_ray* h_rays, size_t n_rays, void* h_out, bool counts, float* ms4)
{
    if (!mg || (!h_rays && n_rays) || (!h_out && n_rays)) return gb_set_error(GRACE_B200_EINVAL, "NULL argument");
    if (n_rays % 32) return gb_set_error(GRACE_B200_EINVAL, "Number of rays must be a multiple of the warp size (32).");
    if (!mg->n_leaves) return gb_set_error(GRACE_B200_EINVAL, "no tree: call grace_b200_mgpu_build_f4 first");
    if (n_rays == 0) return GRACE_B200_OK;
    const int world = (int)mg->d.size();
    Dev& d0 = mg->d[0];
    size_t per = 0;
    for (int r = 0; r < world; ++r) per = std::max(per, local_count(n_rays, r, world));
    cudaSetDevice(d0.id);
    Ev ev;
    int rc;
    if ((rc = ensure(&mg->rays_all, &mg->rays_all_cap, n_rays))) return rc;
    if (mg->res_cap < (size_t)world * per || mg->res_cap < n_rays) {
        if (mg->gathered) cudaFree(mg->gathered);
        if (mg->ordered) cudaFree(mg->ordered);
        mg->gathered = mg->ordered = nullptr;
        mg->res_cap = std::max((size_t)world * per, n_rays);
        MG_CUDA(cudaMalloc((void**)&mg->gathered, mg->res_cap * 4));
        MG_CUDA(cudaMalloc((void**)&mg->ordered, mg->res_cap * 4));
    }
    // ---- rays: host -> device 0 -> every device (ncclBroadcast) -> each keeps its tiles ----
    MG_CUDA(cudaEventRecord(ev.e[0], d0.st));
    MG_CUDA(cudaMemcpyAsync(mg->rays_all, h_rays, n_rays * sizeof(grace_b200_ray), cudaMemcpyHostToDevice, d0.st));
    rc = for_each_device(mg, [&](Dev& dv) -> int {
        int r;
        if (world > 1 && &dv != &mg->d[0] && (r = ensure(&dv.rays_full, &dv.rays_full_cap, n_rays))) return r;
        if (world > 1 && (r = ensure(&dv.rays, &dv.rays_cap, per))) return r;
        return ensure(&dv.out, &dv.out_cap, per);
    });
    if (rc) return rc;
    if (world > 1) {
        MG_NCCL(ncclGroupStart());
        for (int r = 0; r < world; ++r) {
            Dev& dv = mg->d[r];
            MG_NCCL(ncclBroadcast(mg->rays_all, r == 0 ? (void*)mg->rays_all : (void*)dv.rays_full, n_rays * 7, ncclFloat, 0, dv.comm, dv.st));
        }
        MG_NCCL(ncclGroupEnd());
    }
    MG_CUDA(cudaEventRecord(ev.e[1], d0.st));
    // ---- trace: every device its own tiles ----
    rc = for_each_device(mg, [&](Dev& dv) -> int {
        const int rank = (int)(&dv - &mg->d[0]);
        dv.n_local = local_count(n_rays, rank, world);
        const grace_b200_ray* mine = mg->rays_all;
        if (world > 1) {
            take_tiles_kernel<<<128, 256, 0, dv.st>>>((const float*)(rank == 0 ? mg->rays_all : dv.rays_full), (float*)dv.rays, n_rays, rank, world);
            mine = dv.rays;
        }
        grace_b200_tree tr = { dv.nodes, dv.leaves, dv.root, mg->n_leaves, mg->max_per_leaf };
        if (dv.n_local == 0) return GRACE_B200_OK;
        return counts ? grace_b200_trace_hitcounts_f4(dv.ctx, mine, dv.n_local, dv.spheres, mg->n, &tr, (int*)dv.out, dv.st)
                      : grace_b200_trace_cumulative_f4(dv.ctx, mine, dv.n_local, dv.spheres, mg->n, &tr, dv.out, dv.st);
    });
    if (rc) return rc;
    // the slowest device's trace time: wait for all, on device 0's clock
    for (auto& dv : mg->d) { cudaSetDevice(dv.id); MG_CUDA(cudaStreamSynchronize(dv.st)); }
    cudaSetDevice(d0.id);
    MG_CUDA(cudaEventRecord(ev.e[2], d0.st));
    // ---- gather to device 0 (NCCL send/recv), back to ray order, to the host ----
    if (world > 1) {
        MG_NCCL(ncclGroupStart());
        for (int r = 0; r < world; ++r) {
            Dev& dv = mg->d[r];
            if (r == 0) {
                for (int s = 1; s < world; ++s) MG_NCCL(ncclRecv(mg->gathered + (size_t)s * per, per, ncclFloat, s, d0.comm, d0.st));
            } else {
                MG_NCCL(ncclSend(dv.out, per, ncclFloat, 0, dv.comm, dv.st));
            }
        }
        MG_NCCL(ncclGroupEnd());
        MG_CUDA(cudaMemcpyAsync(mg->gathered, d0.out, per * 4, cudaMemcpyDeviceToDevice, d0.st));
        reorder_kernel<<<256, 256, 0, d0.st>>>(mg->gathered, mg->ordered, n_rays, world, per);
    }
    const float* result = world > 1 ? mg->ordered : d0.out;
    MG_CUDA(cudaEventRecord(ev.e[3], d0.st));
    MG_CUDA(cudaMemcpyAsync(h_out, result, n_rays * 4, cudaMemcpyDeviceToHost, d0.st));
    MG_CUDA(cudaEventRecord(ev.e[4], d0.st));
    for (auto& dv : mg->d) { cudaSetDevice(dv.id); MG_CUDA(cudaStreamSynchronize(dv.st)); }
    // a traversal that overflowed its stack or did not terminate leaves short results: never silently
    for (auto& dv : mg->d) {
        int flag = 0;
        cudaSetDevice(dv.id);
        if ((rc = grace_b200_device_error(dv.ctx, &flag, dv.st))) return rc;
        if (flag) return gb_set_error(GRACE_B200_EDEVICE, "device-side traversal error %d on device %d", flag, dv.id);
    }
    cudaSetDevice(d0.id);
    if (ms4) for (int k = 0; k < 4; ++k) cudaEventElapsedTime(ms4 + k, ev.e[k], ev.e[k + 1]);
    return GRACE_B200_OK;
}

} // namespace

extern "C" {

int grace_b200_mgpu_init(grace_b200_mgpu** out, int n_devices, const int* devices)
{
    if (!out) return gb_set_error(GRACE_B200_EINVAL, "NULL argument");
    *out = nullptr;
    int visible = 0;
    MG_CUDA(cudaGetDeviceCount(&visible));
    if (n_devices <= 0) n_devices = visible;
    if (n_devices < 1 || n_devices > visible) return gb_set_error(GRACE_B200_EINVAL, "%d devices asked for, %d visible", n_devices, visible);
    grace_b200_mgpu* mg = new grace_b200_mgpu();
    mg->d.resize(n_devices);
    std::vector<int> ids(n_devices);
    for (int i = 0; i < n_devices; ++i) ids[i] = devices ? devices[i] : i;
    std::vector<ncclComm_t> comms(n_devices);
    MG_NCCL(ncclCommInitAll(comms.data(), n_devices, ids.data()));
    for (int i = 0; i < n_devices; ++i) {
        Dev& dv = mg->d[i];
        dv.id = ids[i];
        dv.comm = comms[i];
        MG_CUDA(cudaSetDevice(dv.id));
        int rc = grace_b200_create(&dv.ctx, dv.id);
        if (rc) return rc;
        MG_CUDA(cudaStreamCreateWithFlags(&dv.st, cudaStreamNonBlocking));
        MG_CUDA(cudaMalloc((void**)&dv.root, sizeof(int)));
    }
    *out = mg;
    return GRACE_B200_OK;
}

int grace_b200_mgpu_finalize(grace_b200_mgpu* mg)
{
    if (!mg) return GRACE_B200_OK;
    for (auto& dv : mg->d) {
        cudaSetDevice(dv.id);
        cudaStreamSynchronize(dv.st);
        if (dv.comm) ncclCommDestroy(dv.comm);
        cudaFree(dv.spheres); cudaFree(dv.nodes); cudaFree(dv.leaves); cudaFree(dv.root); cudaFree(dv.rays); cudaFree(dv.rays_full); cudaFree(dv.out);
        if (dv.st) cudaStreamDestroy(dv.st);
        grace_b200_destroy(dv.ctx);
    }
    if (!mg->d.empty()) { cudaSetDevice(mg->d[0].id); cudaFree(mg->rays_all); cudaFree(mg->gathered); cudaFree(mg->ordered); }
    delete mg;
    return GRACE_B200_OK;
}

int grace_b200_mgpu_n_devices(const grace_b200_mgpu* mg) { return mg ? (int)mg->d.size() : 0; }
const char* grace_b200_mgpu_last_error(void) { return g_mg_err; }

int grace_b200_mgpu_build_f4(grace_b200_mgpu* mg, const float* h_spheres4, size_t n, int max_per_leaf, int key_bits, int how,
                             int* h_n_leaves, float* ms3)
{
    if (!mg || !h_spheres4) return gb_set_error(GRACE_B200_EINVAL, "NULL argument");
    if (key_bits != 30 && key_bits != 63) return gb_set_error(GRACE_B200_EINVAL, "key_bits must be 30 or 63");
    if (how != GRACE_B200_MGPU_BUILD_EVERYWHERE && how != GRACE_B200_MGPU_BUILD_ON_ROOT) return gb_set_error(GRACE_B200_EINVAL, "unknown build mode %d", how);
    const int world = (int)mg->d.size();
    Dev& d0 = mg->d[0];
    mg->n = n; mg->max_per_leaf = max_per_leaf; mg->n_leaves = 0;
    int rc = for_each_device(mg, [&](Dev& dv) -> int {
        if (dv.spheres) cudaFree(dv.spheres);
        dv.spheres = nullptr;
        MG_CUDA(cudaMalloc((void**)&dv.spheres, n * 16));
        size_t cap = dv.leaves_cap;
        int r = ensure(&dv.leaves, &cap, n);          // leaf capacity: n (the count is known after clustering)
        dv.leaves_cap = cap;
        return r;
    });
    if (rc) return rc;
    cudaSetDevice(d0.id);
    Ev ev;
    MG_CUDA(cudaEventRecord(ev.e[0], d0.st));
    MG_CUDA(cudaMemcpyAsync(d0.spheres, h_spheres4, n * 16, cudaMemcpyHostToDevice, d0.st));
    MG_CUDA(cudaEventRecord(ev.e[1], d0.st));
    float t_bcast = 0.f, t_build = 0.f;

    // one device: keys + sort, deltas, leaves (-> count), nodes sized by the count
    auto build_one = [&](Dev& dv, int* L_out) -> int {
        int r;
        if ((r = grace_b200_morton_sort_f4(dv.ctx, dv.spheres, n, key_bits, nullptr, nullptr, nullptr, dv.st))) return r;
        float* deltas = nullptr;
        MG_CUDA(cudaMalloc((void**)&deltas, (n + 1) * 4));
        r = grace_b200_deltas_euclid_f4(dv.ctx, dv.spheres, n, deltas, dv.st);
        int L = 0;
        if (!r) r = grace_b200_albvh_leaves(dv.ctx, deltas, GRACE_B200_DELTA_F32, n, max_per_leaf, dv.leaves, &L, dv.st);
        if (!r) { size_t cap = dv.nodes_cap; r = ensure(&dv.nodes, &cap, 4 * (size_t)(L - 1)); dv.nodes_cap = cap; }
        if (!r) r = grace_b200_albvh_nodes_f4(dv.ctx, dv.spheres, dv.leaves, (size_t)L, nullptr, GRACE_B200_DELTA_F32, dv.nodes, dv.root, dv.st);
        cudaStreamSynchronize(dv.st);
        cudaFree(deltas);
        *L_out = L;
        return r;
    };

    if (how == GRACE_B200_MGPU_BUILD_EVERYWHERE) {
        if (world > 1) {
            MG_NCCL(ncclGroupStart());
            for (auto& dv : mg->d) MG_NCCL(ncclBroadcast(d0.spheres, dv.spheres, n * 4, ncclFloat, 0, dv.comm, dv.st));
            MG_NCCL(ncclGroupEnd());
        }
        MG_CUDA(cudaEventRecord(ev.e[2], d0.st));
        std::vector<int> Ls(world, 0);
        rc = for_each_device(mg, [&](Dev& dv) -> int { return build_one(dv, &Ls[&dv - &mg->d[0]]); });
        if (rc) return rc;
        for (int r = 1; r < world; ++r)
            if (Ls[r] != Ls[0]) return gb_set_error(GRACE_B200_ECUDA, "devices built different trees (%d vs %d leaves)", Ls[r], Ls[0]);
        mg->n_leaves = Ls[0];
        cudaSetDevice(d0.id);
        MG_CUDA(cudaEventRecord(ev.e[3], d0.st));
        MG_CUDA(cudaEventSynchronize(ev.e[3]));
        cudaEventElapsedTime(&t_bcast, ev.e[1], ev.e[2]);
        cudaEventElapsedTime(&t_build, ev.e[2], ev.e[3]);
    } else {
        int L = 0;
        cudaSetDevice(d0.id);
        if ((rc = build_one(d0, &L))) return rc;
        mg->n_leaves = L;
        MG_CUDA(cudaEventRecord(ev.e[2], d0.st));
        if (world > 1) {
            rc = for_each_device(mg, [&](Dev& dv) -> int {
                if (&dv == &mg->d[0]) return GRACE_B200_OK;
                size_t cap = dv.nodes_cap;
                int r = ensure(&dv.nodes, &cap, 4 * (size_t)(L - 1));
                dv.nodes_cap = cap;
                return r;
            });
            if (rc) return rc;
            MG_NCCL(ncclGroupStart());
            for (auto& dv : mg->d) {
                MG_NCCL(ncclBroadcast(d0.spheres, dv.spheres, n * 4, ncclFloat, 0, dv.comm, dv.st));
                MG_NCCL(ncclBroadcast(d0.nodes, dv.nodes, 16 * (size_t)(L - 1), ncclInt32, 0, dv.comm, dv.st));
                MG_NCCL(ncclBroadcast(d0.leaves, dv.leaves, 4 * (size_t)L, ncclInt32, 0, dv.comm, dv.st));
                MG_NCCL(ncclBroadcast(d0.root, dv.root, 1, ncclInt32, 0, dv.comm, dv.st));
            }
            MG_NCCL(ncclGroupEnd());
        }
        cudaSetDevice(d0.id);
        MG_CUDA(cudaEventRecord(ev.e[3], d0.st));
        for (auto& dv : mg->d) { cudaSetDevice(dv.id); MG_CUDA(cudaStreamSynchronize(dv.st)); }
        cudaEventElapsedTime(&t_build, ev.e[1], ev.e[2]);
        cudaEventElapsedTime(&t_bcast, ev.e[2], ev.e[3]);
    }
    if (h_n_leaves) *h_n_leaves = mg->n_leaves;
    if (ms3) { cudaEventElapsedTime(ms3, ev.e[0], ev.e[1]); ms3[1] = t_bcast; ms3[2] = t_build; }
    return GRACE_B200_OK;
}

int grace_b200_mgpu_trace_cumulative_f4(grace_b200_mgpu* mg, const grace_b200_ray* h_rays, size_t n_rays, float* h_cumulated, float* ms4)
{
    return trace_any(mg, h_rays, n_rays, h_cumulated, false, ms4);
}

int grace_b200_mgpu_trace_hitcounts_f4(grace_b200_mgpu* mg, const grace_b200_ray* h_rays, size_t n_rays, int* h_hit_counts, float* ms4)
{
    return trace_any(mg, h_rays, n_rays, h_hit_counts, true, ms4);
}

int grace_b200_mgpu_copy_tree(grace_b200_mgpu* mg, int dev, float* h_spheres4, int* h_nodes16, int* h_leaves4, int* h_root)
{
    if (!mg || dev < 0 || dev >= (int)mg->d.size() || !mg->n_leaves) return gb_set_error(GRACE_B200_EINVAL, "bad argument");
    Dev& dv = mg->d[dev];
    MG_CUDA(cudaSetDevice(dv.id));
    MG_CUDA(cudaStreamSynchronize(dv.st));
    if (h_spheres4) MG_CUDA(cudaMemcpy(h_spheres4, dv.spheres, mg->n * 16, cudaMemcpyDeviceToHost));
    if (h_nodes16) MG_CUDA(cudaMemcpy(h_nodes16, dv.nodes, 64 * (size_t)(mg->n_leaves - 1), cudaMemcpyDeviceToHost));
    if (h_leaves4) MG_CUDA(cudaMemcpy(h_leaves4, dv.leaves, 16 * (size_t)mg->n_leaves, cudaMemcpyDeviceToHost));
    if (h_root) MG_CUDA(cudaMemcpy(h_root, dv.root, sizeof(int), cudaMemcpyDeviceToHost));
    return GRACE_B200_OK;
}

} // extern "C"
