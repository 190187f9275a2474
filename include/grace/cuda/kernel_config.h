// grace/cuda/kernel_config.h -- launch constants user kernels written against GRACE refer to
// (reference: include/grace/cuda/kernel_config.h:5-13).  The library's own kernels size their grids
// from the device (148 SMs x resident CTAs); MAX_BLOCKS is only a cap for user-side grid-stride
// launches, so it is set for a B200 (148 SMs x 16) where the reference has 7 Kepler SMX x 16.
#pragma once

namespace grace {

const int MORTON_THREADS_PER_BLOCK = 512;
const int BUILD_THREADS_PER_BLOCK = 512;
const int SHIFTS_THREADS_PER_BLOCK = 512;
const int AABB_THREADS_PER_BLOCK = 512;
const int TRACE_THREADS_PER_BLOCK = 256;
const int RAYS_THREADS_PER_BLOCK = 512;
const int MAX_BLOCKS = 148 * 16;
const int WARP_SIZE = 32;
const int STACK_SIZE = 64;

} // namespace grace
