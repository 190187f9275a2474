// radix_sort.cuh -- internal interface of the onesweep key/index radix sort.
#pragma once
#include "common.cuh"

// Bytes of workspace needed by gb_sort_pairs for n keys of key_bytes each.
size_t gb_sort_workspace_bytes(size_t n, int key_bytes);

// Stable LSD radix sort of (key, index) pairs on bits [0, key_bits).
//   d_keys_in   : n keys (not modified unless it aliases d_keys_out)
//   d_keys_out  : n sorted keys (may alias d_keys_in)
//   d_perm_out  : n indices, sorted[i] = in[perm[i]]
//   ws          : gb_sort_workspace_bytes() bytes of scratch
// If d_hist is non-NULL it already holds the per-pass digit histograms
// ([passes][256] uint32, produced by a fused producer kernel) and the histogram
// pass over the keys is skipped.
// d_rec16_in / d_rec16_out (both or neither; more than one pass): the last pass gathers the 16-byte
// records, rec_out[i] = rec_in[perm[i]], instead of writing the permutation (d_perm_out is not
// written), and writes the sorted keys only if d_keys_out is not NULL.
template <typename KeyT>
int gb_sort_pairs(grace_b200_ctx* ctx, const KeyT* d_keys_in, KeyT* d_keys_out,
                  uint32_t* d_perm_out, size_t n, int key_bits, void* ws,
                  const uint32_t* d_hist, cudaStream_t st, const void* d_rec16_in = nullptr,
                  void* d_rec16_out = nullptr);

// out[i] = in[perm[i]] for records of rec_bytes (4, 16 or 28).
int gb_gather_records(const void* d_in, void* d_out, const uint32_t* d_perm, size_t n,
                      int rec_bytes, int sm_count, cudaStream_t st);

constexpr int GB_RADIX_BITS = 8;
constexpr int GB_RADIX = 1 << GB_RADIX_BITS;
