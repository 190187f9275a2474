// grace/generic/meta.h -- type maps between vector and scalar reals (reference: generic/meta.h:14-75).
#pragma once
#include "grace/types.h"

namespace grace {

template <typename> struct Real3ToRealMapper;
template <> struct Real3ToRealMapper<float3> { typedef float type; };
template <> struct Real3ToRealMapper<double3> { typedef double type; };

template <typename> struct Real4ToRealMapper;
template <> struct Real4ToRealMapper<float4> { typedef float type; };
template <> struct Real4ToRealMapper<double4> { typedef double type; };

template <typename> struct RealToReal3Mapper;
template <> struct RealToReal3Mapper<float> { typedef float3 type; };
template <> struct RealToReal3Mapper<double> { typedef double3 type; };

template <typename> struct RealToReal4Mapper;
template <> struct RealToReal4Mapper<float> { typedef float4 type; };
template <> struct RealToReal4Mapper<double> { typedef double4 type; };

} // namespace grace
