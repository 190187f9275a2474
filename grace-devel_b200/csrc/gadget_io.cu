// gadget_io.cu -- Gadget-2 (type 1) snapshot loader: the step BEFORE the hot path
// (SURVEY.md 8f N1).
//
// Reference behaviour (GRACE): read_gadget, tests/helper/read_gadget.cuh:69-167 -- one
// ifstream::read per 4-byte field into a host_vector<float4> (positions of the gas particles,
// then their smoothing lengths from the HSML block), every other block skipped 4 bytes at a
// time, followed by a blocking host_vector -> device_vector copy (:161-167).  File layout as
// that reader assumes it: 256-byte header {int npart[6]; double mass[6]; fill} and the blocks
// POS (3 floats x all particles, types in order), VEL (3), ID (1 x 4 bytes), MASS (1, only the
// types whose header mass is 0; the block exists only if there is such a particle), U, RHO,
// HSML (1 x gas particles each), every block wrapped in 4-byte markers.
//
// B200 design: the block offsets follow from the header alone, so nothing is skipped by
// reading.  The gas positions and smoothing lengths are pread() in large chunks into two
// pinned staging buffers; while the copy engine moves chunk c to the device (cudaMemcpyAsync)
// the host reads chunk c+1; a small kernel interleaves {x,y,z} and h into the float4 records
// the build expects.  Everything is ordered on the caller's stream, so the key/sort stages
// queued behind the call start as the last chunk lands, with no host synchronisation after the
// final read.
#include "common.cuh"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

namespace {

struct GadgetLayout {
    long long npart[6];
    double mass[6];
    long long n_total, n_gas, n_withmass;
    long long pos_off, hsml_off, file_bytes_needed;
};

// File offsets of the data read_gadget.cuh reads, from the header alone.
int gadget_layout(int fd, GadgetLayout* L)
{
    unsigned char hdr[4 + 256 + 4];
    const ssize_t got = pread(fd, hdr, sizeof(hdr), 0);
    if (got != (ssize_t)sizeof(hdr)) return gb_set_error(GRACE_B200_EINVAL, "Gadget file shorter than its header");
    int np[6];
    memcpy(np, hdr + 4, sizeof(np));
    memcpy(L->mass, hdr + 4 + sizeof(np), sizeof(L->mass));
    L->n_total = L->n_withmass = 0;
    for (int i = 0; i < 6; ++i) {
        if (np[i] < 0) return gb_set_error(GRACE_B200_EINVAL, "negative particle count in Gadget header");
        L->npart[i] = np[i];
        L->n_total += np[i];
        if (L->mass[i] == 0) L->n_withmass += np[i];       // read_gadget.cuh:99-101
    }
    L->n_gas = L->npart[0];
    long long off = sizeof(hdr);
    L->pos_off = off + 4;                                  // gas comes first inside every block
    off += 4 + 12 * L->n_total + 4;                        // POS
    off += 4 + 12 * L->n_total + 4;                        // VEL   (:123)
    off += 4 + 4 * L->n_total + 4;                         // ID    (:126)
    if (L->n_withmass > 0) off += 4 + 4 * L->n_withmass + 4;   // MASS  (:131-141)
    off += 4 + 4 * L->n_gas + 4;                           // U     (:147-149)
    off += 4 + 4 * L->n_gas + 4;                           // RHO   (:152-154)
    L->hsml_off = off + 4;                                 // HSML  (:157-161)
    L->file_bytes_needed = L->hsml_off + 4 * L->n_gas;
    return GRACE_B200_OK;
}

int read_fully(int fd, void* dst, size_t bytes, long long off)
{
    char* p = (char*)dst;
    while (bytes) {
        const ssize_t got = pread(fd, p, bytes, off);
        if (got < 0) { if (errno == EINTR) continue; return gb_set_error(GRACE_B200_EINVAL, "read: %s", strerror(errno)); }
        if (got == 0) return gb_set_error(GRACE_B200_EINVAL, "Gadget file truncated");
        p += got; off += got; bytes -= (size_t)got;
    }
    return GRACE_B200_OK;
}

// pos: 3 floats per particle, hsml: 1 float per particle -> float4 {x, y, z, h}
__global__ void __launch_bounds__(256)
interleave_kernel(const float* __restrict__ pos, const float* __restrict__ hsml, size_t n, float4* __restrict__ out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], hsml[i]);
}

constexpr size_t CHUNK = 1u << 21;     // particles per staging chunk: 24 MiB positions + 8 MiB h

// Pinned staging buffers live in the context (pinning 64 MiB costs as much as reading it).
int ensure_staging(grace_b200_ctx* ctx, size_t bytes)
{
    if (ctx->stage_bytes >= bytes) return GRACE_B200_OK;
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage_done[i]) GB_CUDA(cudaEventSynchronize(ctx->stage_done[i]));
        if (ctx->stage_host[i]) { cudaFreeHost(ctx->stage_host[i]); ctx->stage_host[i] = nullptr; }
    }
    ctx->stage_bytes = 0;
    for (int i = 0; i < 2; ++i) {
        GB_CUDA(cudaMallocHost((void**)&ctx->stage_host[i], bytes));
        if (!ctx->stage_done[i]) GB_CUDA(cudaEventCreateWithFlags(&ctx->stage_done[i], cudaEventDisableTiming));
    }
    ctx->stage_bytes = bytes;
    return GRACE_B200_OK;
}

} // namespace

extern "C" {

int grace_b200_gadget_info(const char* path, long long* npart6, double* mass6, long long* n_gas)
{
    GB_REQUIRE(path, GRACE_B200_EINVAL, "NULL path");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return gb_set_error(GRACE_B200_EINVAL, "cannot open %s: %s", path, strerror(errno));
    GadgetLayout L;
    const int rc = gadget_layout(fd, &L);
    close(fd);
    if (rc) return rc;
    for (int i = 0; i < 6; ++i) {
        if (npart6) npart6[i] = L.npart[i];
        if (mass6) mass6[i] = L.mass[i];
    }
    if (n_gas) *n_gas = L.n_gas;
    return GRACE_B200_OK;
}

int grace_b200_read_gadget_f4(grace_b200_ctx* ctx, const char* path, float* d_spheres4, size_t capacity,
                              size_t* n_gas_out, void* stream)
{
    GB_REQUIRE(ctx && path && d_spheres4, GRACE_B200_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return gb_set_error(GRACE_B200_EINVAL, "cannot open %s: %s", path, strerror(errno));
    struct Closer { int fd; ~Closer() { close(fd); } } closer{ fd };
    GadgetLayout L;
    int rc = gadget_layout(fd, &L);
    if (rc) return rc;
    // read_gadget.cuh:85-90
    GB_REQUIRE(L.n_gas > 0, GRACE_B200_EINVAL, "Gadget file %s has no gas particles!", path);
    struct stat sb;
    if (fstat(fd, &sb) == 0 && (long long)sb.st_size < L.file_bytes_needed)
        return gb_set_error(GRACE_B200_EINVAL, "Gadget file %s is truncated (%lld bytes, %lld needed)", path,
                            (long long)sb.st_size, L.file_bytes_needed);
    const size_t n = (size_t)L.n_gas;
    if (n_gas_out) *n_gas_out = n;
    GB_REQUIRE(capacity >= n, GRACE_B200_ERANGE, "%zu gas particles do not fit in a buffer of %zu", n, capacity);

    const size_t chunk = std::min(n, CHUNK);
    const size_t buf_bytes = gb_align(chunk * 16);        // [positions 12 B | h 4 B] per particle
    rc = ensure_staging(ctx, buf_bytes);
    if (rc) return rc;
    char* d_stage = (char*)gb_workspace(ctx, 2 * buf_bytes);       // device twins of the two buffers
    if (!d_stage) return GRACE_B200_ENOMEM;
    int which = 0;
    for (size_t first = 0; first < n; first += chunk, which ^= 1) {
        const size_t m = std::min(chunk, n - first);
        char* h_buf = ctx->stage_host[which];
        char* d_buf = d_stage + which * buf_bytes;
        // the previous H2D copy out of this pinned buffer (and the kernel reading its device twin)
        // must have finished before the host overwrites it
        GB_CUDA(cudaEventSynchronize(ctx->stage_done[which]));
        rc = read_fully(fd, h_buf, m * 12, L.pos_off + (long long)first * 12);
        if (rc) return rc;
        rc = read_fully(fd, h_buf + m * 12, m * 4, L.hsml_off + (long long)first * 4);
        if (rc) return rc;
        GB_CUDA(cudaMemcpyAsync(d_buf, h_buf, m * 16, cudaMemcpyHostToDevice, st));
        const int blocks = (int)std::min<size_t>((m + 255) / 256, (size_t)ctx->sm_count * 8);
        interleave_kernel<<<blocks, 256, 0, st>>>((const float*)d_buf, (const float*)(d_buf + m * 12), m,
                                                   (float4*)d_spheres4 + first);
        GB_LAUNCH_CHECK();
        GB_CUDA(cudaEventRecord(ctx->stage_done[which], st));
    }
    // No synchronisation here: the interleave kernels are ordered before whatever the caller
    // queues on the stream next (the workspace twins are reused only by later calls on it).
    return GRACE_B200_OK;
}

int grace_b200_write_gadget_f4(const char* path, const float* h_spheres4, size_t n_gas, size_t n_other,
                               int other_has_mass_block)
{
    GB_REQUIRE(path && (h_spheres4 || n_gas == 0), GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(n_gas < (1ull << 31) && n_other < (1ull << 31), GRACE_B200_ERANGE, "Gadget counts are 32-bit");
    FILE* f = fopen(path, "wb");
    if (!f) return gb_set_error(GRACE_B200_EINVAL, "cannot create %s: %s", path, strerror(errno));
    const size_t n_total = n_gas + n_other;
    auto marker = [&](size_t bytes) { const int m = (int)bytes; fwrite(&m, 4, 1, f); };
    unsigned char hdr[256];
    memset(hdr, 0, sizeof(hdr));
    int np[6] = { (int)n_gas, (int)n_other, 0, 0, 0, 0 };
    double mass[6] = { 1.0, other_has_mass_block ? 0.0 : 2.0, 0, 0, 0, 0 };     // mass 0 => per-particle MASS block
    memcpy(hdr, np, sizeof(np));
    memcpy(hdr + sizeof(np), mass, sizeof(mass));
    marker(256); fwrite(hdr, 256, 1, f); marker(256);
    const size_t B = 1 << 16;
    float* tmp = (float*)malloc(B * 3 * sizeof(float));
    if (!tmp) { fclose(f); return gb_set_error(GRACE_B200_ENOMEM, "out of host memory"); }
    auto block3 = [&](bool positions) {          // POS / VEL
        marker(n_total * 12);
        for (size_t b = 0; b < n_total; b += B) {
            const size_t m = std::min(B, n_total - b);
            for (size_t i = 0; i < m; ++i) {
                const size_t p = b + i;
                for (int k = 0; k < 3; ++k)
                    tmp[3 * i + k] = positions ? (p < n_gas ? h_spheres4[4 * p + k] : 0.25f + 0.001f * (float)k) : 0.0f;
            }
            fwrite(tmp, 12, m, f);
        }
        marker(n_total * 12);
    };
    auto block1 = [&](size_t count, int what) {   // ID, MASS, U, RHO, HSML
        marker(count * 4);
        for (size_t b = 0; b < count; b += B) {
            const size_t m = std::min(B, count - b);
            for (size_t i = 0; i < m; ++i) {
                const size_t p = b + i;
                if (what == 0) { const int id = (int)p; memcpy(tmp + i, &id, 4); }
                else if (what == 4) tmp[i] = h_spheres4[4 * p + 3];
                else tmp[i] = 1.0f;
            }
            fwrite(tmp, 4, m, f);
        }
        marker(count * 4);
    };
    block3(true);
    block3(false);
    block1(n_total, 0);
    if (other_has_mass_block && n_other > 0) block1(n_other, 1);
    block1(n_gas, 2);
    block1(n_gas, 3);
    block1(n_gas, 4);
    free(tmp);
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) return gb_set_error(GRACE_B200_EINVAL, "write to %s failed", path);
    return GRACE_B200_OK;
}

} // extern "C"
