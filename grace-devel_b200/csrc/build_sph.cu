// build_sph.cu -- the fused "keys + sort" entry point of the SPH build path.
//
// Replaces morton_keys30_sort_sph / morton_keys63_sort_sph (GRACE cuda/build_sph.cuh:41-82):
// there, a temporary key vector is allocated per call, bounds cost two Thrust reductions
// with host read-backs, and thrust::sort_by_key moves the 16-byte spheres through every
// radix pass.  Here: bounds (optional) -> keys -> onesweep on (key, index), the last pass
// gathering the spheres, all on the caller's stream with no host synchronisation.
#include "common.cuh"
#include "radix_sort.cuh"

template <typename KeyT>
int gb_launch_morton_keys(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                          const float* d_bounds6, const float* h_bounds6, KeyT* d_keys,
                          cudaStream_t st);

template <typename KeyT>
int gb_launch_morton_keys_fused(grace_b200_ctx* ctx, const float* d_spheres4, size_t n,
                                const float* d_bounds6, const float* h_bounds6, KeyT* d_keys,
                                uint32_t* d_hist, int passes, float* d_copy4, cudaStream_t st);

namespace {

template <typename KeyT>
int morton_sort_typed(grace_b200_ctx* ctx, float* d_spheres4, size_t n, int key_bits,
                      const float* h_bot3, const float* h_top3, void* d_keys_out, cudaStream_t st)
{
    const size_t sort_ws = gb_sort_workspace_bytes(n, (int)sizeof(KeyT));
    const size_t bytes = sort_ws + gb_align(n * sizeof(KeyT)) + gb_align(n * 4) +
                         gb_align(n * 16) + 512 + 8 * 256 * 4;
    char* ws = (char*)gb_workspace(ctx, bytes);
    if (!ws) return GRACE_B200_ENOMEM;
    char* p = ws + sort_ws;
    KeyT* keys = (KeyT*)p;            p += gb_align(n * sizeof(KeyT));
    uint32_t* perm = (uint32_t*)p;    p += gb_align(n * 4);
    float4* tmp = (float4*)p;         p += gb_align(n * 16);
    float* d_bounds = (float*)p;      p += 512;
    uint32_t* hist = (uint32_t*)p;
    const int sort_bits = key_bits == 30 ? 32 : 64;
    const int passes = sort_bits / 8;
    GB_CUDA(cudaMemsetAsync(hist, 0, (size_t)passes * 256 * 4, st));
    int rc;
    // one pass over the spheres: keys, the digit counts of every radix pass, and the copy the
    // final gather reads (so the gather writes the caller's array directly)
    if (h_bot3 && h_top3) {
        const float hb[6] = { h_bot3[0], h_bot3[1], h_bot3[2], h_top3[0], h_top3[1], h_top3[2] };
        rc = gb_launch_morton_keys_fused<KeyT>(ctx, d_spheres4, n, nullptr, hb, keys, hist, passes, (float*)tmp, st);
    } else {
        rc = grace_b200_bounds_f4(ctx, d_spheres4, n, d_bounds, st);
        if (rc) return rc;
        // bounds kernel used the head of the arena for its partials; they are consumed
        // before the sort touches the same bytes (stream order).
        rc = gb_launch_morton_keys_fused<KeyT>(ctx, d_spheres4, n, d_bounds, nullptr, keys, hist, passes, (float*)tmp, st);
    }
    if (rc) return rc;
    // the last radix pass writes the spheres to their sorted places (no permutation, no separate gather launch)
    rc = gb_sort_pairs<KeyT>(ctx, keys, d_keys_out ? keys : nullptr, perm, n, sort_bits, ws, hist, st, tmp, d_spheres4);
    if (rc) return rc;
    if (d_keys_out)
        GB_CUDA(cudaMemcpyAsync(d_keys_out, keys, n * sizeof(KeyT), cudaMemcpyDeviceToDevice, st));
    return GRACE_B200_OK;
}

} // namespace

extern "C" int grace_b200_morton_sort_f4(grace_b200_ctx* ctx, float* d_spheres4, size_t n,
                                         int key_bits, const float* h_bot3, const float* h_top3,
                                         void* d_keys_out, void* stream)
{
    GB_REQUIRE(ctx && d_spheres4, GRACE_B200_EINVAL, "NULL argument");
    GB_REQUIRE(key_bits == 30 || key_bits == 63, GRACE_B200_EINVAL, "key_bits must be 30 or 63");
    GB_REQUIRE((h_bot3 == nullptr) == (h_top3 == nullptr), GRACE_B200_EINVAL,
               "give both bounds or neither");
    if (n == 0) return GRACE_B200_OK;
    if (key_bits == 30)
        return morton_sort_typed<uint32_t>(ctx, d_spheres4, n, key_bits, h_bot3, h_top3, d_keys_out,
                                           (cudaStream_t)stream);
    return morton_sort_typed<uint64_t>(ctx, d_spheres4, n, key_bits, h_bot3, h_top3, d_keys_out,
                                       (cudaStream_t)stream);
}
