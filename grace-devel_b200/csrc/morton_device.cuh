// morton_device.cuh -- device-side Morton key arithmetic.
//
// Restates (not copies) GRACE's key definition so results are bit-identical:
//   generic/bits.h:24-46      bit spreading (10 bits -> 30, 21 bits -> 63)
//   generic/morton.h:14-29    key = spread(z)<<2 | spread(y)<<1 | spread(x)
//   cuda/kernels/morton.cuh:46-48   q = (KeyT)(scale * (c - bot))   FADD, FMUL, F2I.TRUNC
//   cuda/kernels/morton.cuh:107-113 scale = span / (top - bot), float division
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

template <typename KeyT> struct GbKeyTraits;
template <> struct GbKeyTraits<uint32_t> { static constexpr int span = (1 << 10) - 1; };
template <> struct GbKeyTraits<uint64_t> { static constexpr int span = (1 << 21) - 1; };

__device__ __forceinline__ uint32_t gb_spread10(uint32_t x)
{
    x &= 0x3FFu;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x <<  8)) & 0x0300F00Fu;
    x = (x | (x <<  4)) & 0x030C30C3u;
    x = (x | (x <<  2)) & 0x09249249u;
    return x;
}

__device__ __forceinline__ uint64_t gb_spread21(uint64_t x)
{
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x001f00000000ffffull;
    x = (x | x << 16) & 0x001f0000ff0000ffull;
    x = (x | x <<  8) & 0x100f00f00f00f00full;
    x = (x | x <<  4) & 0x10c30c30c30c30c3ull;
    x = (x | x <<  2) & 0x1249249249249249ull;
    return x;
}

__device__ __forceinline__ uint32_t gb_interleave(uint32_t x, uint32_t y, uint32_t z)
{
    return gb_spread10(z) << 2 | gb_spread10(y) << 1 | gb_spread10(x);
}
__device__ __forceinline__ uint64_t gb_interleave(uint64_t x, uint64_t y, uint64_t z)
{
    return gb_spread21(z) << 2 | gb_spread21(y) << 1 | gb_spread21(x);
}

template <typename KeyT>
__device__ __forceinline__ float3 gb_morton_scale(float3 bot, float3 top)
{
    const float span = (float)GbKeyTraits<KeyT>::span;
    return make_float3(__fdiv_rn(span, __fsub_rn(top.x, bot.x)),
                       __fdiv_rn(span, __fsub_rn(top.y, bot.y)),
                       __fdiv_rn(span, __fsub_rn(top.z, bot.z)));
}

// cvt.rzi saturates and maps NaN to 0, exactly like the reference's static_cast.
__device__ __forceinline__ void gb_quantise(float v, uint32_t& q) { q = __float2uint_rz(v); }
__device__ __forceinline__ void gb_quantise(float v, uint64_t& q) { q = __float2ull_rz(v); }

template <typename KeyT>
__device__ __forceinline__ KeyT gb_morton_key(float4 c, float3 bot, float3 scale)
{
    KeyT x, y, z;
    gb_quantise(__fmul_rn(scale.x, __fsub_rn(c.x, bot.x)), x);
    gb_quantise(__fmul_rn(scale.y, __fsub_rn(c.y, bot.y)), y);
    gb_quantise(__fmul_rn(scale.z, __fsub_rn(c.z, bot.z)), z);
    return gb_interleave(x, y, z);
}
