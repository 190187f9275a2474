// grace/types.h -- base types of the GRACE API (reference: include/grace/types.h:14-51).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>   // float3/float4/int4 and the runtime API used by the shim

#ifdef __CUDACC__
#define GRACE_HOST __host__ inline
#define GRACE_DEVICE __device__ inline
#define GRACE_HOST_DEVICE __host__ __device__ inline
#else
#define GRACE_HOST inline
#define GRACE_HOST_DEVICE inline
#endif

namespace grace {

typedef uint32_t uinteger32;
typedef uint64_t uinteger64;
typedef int32_t integer32;
typedef int64_t integer64;

// +ve = 1, -ve = 0 per axis (x is the high bit): octants for the single-octant ray generator.
enum Octants { MMM = 0, MMP = 1, MPM = 2, MPP = 3, PMM = 4, PMP = 5, PPM = 6, PPP = 7 };

enum RaySortType { NoSort, DirectionSort, EndPointSort };

} // namespace grace
