// grace/error.h -- error conventions of the reference (include/grace/error.h:35-64) on top of
// the C ABI's status codes: argument errors throw std::invalid_argument
// (bintree_trace.cuh:231-238, albvh.cuh:795-799, cuda/gen_rays.cuh:124-130), CUDA failures
// print to stderr and exit.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "grace_b200.h"

namespace grace {
namespace detail {

inline void check(int rc, const char* file, int line)
{
    if (rc == GRACE_B200_OK) return;
    const std::string msg = grace_b200_last_error();
    if (rc == GRACE_B200_EINVAL) throw std::invalid_argument(msg);
    if (rc == GRACE_B200_ERANGE) throw std::length_error(msg);
    if (rc == GRACE_B200_EDEVICE) throw std::runtime_error(msg);
    std::fprintf(stderr, "**** GRACE error in %s at line %d:\n%s\n", file, line, msg.c_str());
    std::exit(rc);
}

inline void cuda_check(cudaError_t e, const char* file, int line)
{
    if (e == cudaSuccess) return;
    std::fprintf(stderr, "**** GRACE CUDA error in %s at line %d:\n%s\n", file, line, cudaGetErrorString(e));
    std::exit((int)e);
}

} // namespace detail
} // namespace grace

#define GRACE_B200_CHECK(call) ::grace::detail::check((call), __FILE__, __LINE__)
#define GRACE_CUDA_CHECK(call) ::grace::detail::cuda_check((call), __FILE__, __LINE__)

// Names user code written against the reference uses directly (error.h:35-64).
namespace grace {
inline void cuda_error_check(cudaError_t code, const char* file, int line, bool terminate = true)
{
    if (code == cudaSuccess) return;
    std::fprintf(stderr, "**** GRACE CUDA Error ****\nFile:  %s\nLine:  %d\nError: %s\n", file, line, cudaGetErrorString(code));
    if (terminate) std::exit((int)code);
}
inline void cuda_kernel_check(const char* file, int line, bool terminate = true)
{
    cuda_error_check(cudaPeekAtLastError(), file, line, terminate);
#ifdef GRACE_DEBUG
    cuda_error_check(cudaDeviceSynchronize(), file, line, terminate);
#endif
}
} // namespace grace
#define GRACE_KERNEL_CHECK() { ::grace::cuda_kernel_check(__FILE__, __LINE__); }
#define GRACE_GOT_TO() std::fprintf(stderr, "At %s@%d\n", __FILE__, __LINE__);
#if defined(GRACE_DEBUG)
#define GRACE_STATIC_ASSERT(predicate, msg) { static_assert(predicate, msg); }
#else
#define GRACE_STATIC_ASSERT(ignore, msg)
#endif
