// grace/device_vector.h -- minimal owning device buffer used where the reference uses
// thrust::device_vector (this repo has no Thrust dependency).  Every shim entry point is a
// template over the container, needing only size() / data() / resize(), so
// thrust::device_vector works unchanged in user code that includes Thrust itself.
#pragma once
#include <cstddef>
#include <type_traits>
#include <utility>
#include <vector>

#include "grace/error.h"
#include "grace/types.h"

namespace grace {

template <typename T>
class device_vector {
public:
    typedef T value_type;
    struct lazy_t {};
    device_vector() : ptr_(nullptr), size_(0), cap_(0) {}
    // n elements whose storage is allocated at the first data() / resize(): grace::Tree is constructed for
    // N leaves (1 GiB of nodes at 2^24) and shrunk to a twentieth of that by the builder
    device_vector(size_t n, lazy_t) : ptr_(nullptr), size_(n), cap_(0) {}
    explicit device_vector(size_t n) : ptr_(nullptr), size_(0), cap_(0) { resize(n); }
    device_vector(size_t n, const T& v) : ptr_(nullptr), size_(0), cap_(0) { resize(n, v); }
    device_vector(const std::vector<T>& h) : ptr_(nullptr), size_(0), cap_(0) { *this = h; }
    device_vector(const device_vector& o) : ptr_(nullptr), size_(0), cap_(0)
    {
        resize(o.size_);
        if (size_ && o.ptr_) GRACE_CUDA_CHECK(cudaMemcpy(ptr_, o.ptr_, size_ * sizeof(T), cudaMemcpyDeviceToDevice));
    }
    device_vector& operator=(const device_vector& o)
    {
        if (this != &o) {
            resize(o.size_);
            if (size_ && o.ptr_) GRACE_CUDA_CHECK(cudaMemcpy(ptr_, o.ptr_, size_ * sizeof(T), cudaMemcpyDeviceToDevice));
        }
        return *this;
    }
    device_vector& operator=(const std::vector<T>& h)
    {
        resize(h.size());
        if (size_) GRACE_CUDA_CHECK(cudaMemcpy(ptr_, h.data(), size_ * sizeof(T), cudaMemcpyHostToDevice));
        return *this;
    }
    ~device_vector() { if (ptr_) cudaFree(ptr_); }

    size_t size() const { return size_; }
    bool empty() const { return size_ == 0; }
    T* data() { materialise(); return ptr_; }
    const T* data() const { const_cast<device_vector*>(this)->materialise(); return ptr_; }

    // Contents are preserved up to min(old, new) elements.
    void resize(size_t n)
    {
        if (n > cap_) {
            T* p = nullptr;
            GRACE_CUDA_CHECK(cudaMalloc((void**)&p, n * sizeof(T)));
            if (size_ && ptr_) GRACE_CUDA_CHECK(cudaMemcpy(p, ptr_, size_ * sizeof(T), cudaMemcpyDeviceToDevice));
            if (ptr_) GRACE_CUDA_CHECK(cudaFree(ptr_));
            ptr_ = p;
            cap_ = n;
        }
        size_ = n;
    }
    void resize(size_t n, const T& v)
    {
        const size_t old = size_;
        resize(n);
        if (n > old) {
            std::vector<T> fill(n - old, v);
            GRACE_CUDA_CHECK(cudaMemcpy(ptr_ + old, fill.data(), (n - old) * sizeof(T), cudaMemcpyHostToDevice));
        }
    }
    void shrink_to_fit()
    {
        if (size_ == cap_) return;
        T* p = nullptr;
        if (size_) {
            GRACE_CUDA_CHECK(cudaMalloc((void**)&p, size_ * sizeof(T)));
            GRACE_CUDA_CHECK(cudaMemcpy(p, ptr_, size_ * sizeof(T), cudaMemcpyDeviceToDevice));
        }
        if (ptr_) GRACE_CUDA_CHECK(cudaFree(ptr_));
        ptr_ = p;
        cap_ = size_;
    }
    std::vector<T> to_host() const
    {
        std::vector<T> h(size_);
        if (size_) GRACE_CUDA_CHECK(cudaMemcpy(h.data(), ptr_, size_ * sizeof(T), cudaMemcpyDeviceToHost));
        return h;
    }
    T operator[](size_t i) const    // blocking element read, like thrust's device_reference
    {
        T v;
        GRACE_CUDA_CHECK(cudaMemcpy(&v, ptr_ + i, sizeof(T), cudaMemcpyDeviceToHost));
        return v;
    }

private:
    void materialise()
    {
        if (ptr_ || size_ == 0) return;
        GRACE_CUDA_CHECK(cudaMalloc((void**)&ptr_, size_ * sizeof(T)));
        cap_ = size_;
    }
    T* ptr_;
    size_t size_, cap_;
};

namespace detail {
// raw pointer of whatever a container's data() returns: T* or thrust::device_ptr<T>
template <typename T> inline T* raw(T* p) { return p; }
template <typename P> inline auto raw(P p) -> decltype(p.get()) { return p.get(); }

// element type of a container (through data()), and SFINAE on it: the float4 entry points go
// through the C ABI, the double4 ones through the header templates
template <typename Vec>
struct elem_of {
    typedef typename std::remove_cv<typename std::remove_pointer<decltype(raw(std::declval<Vec&>().data()))>::type>::type type;
};
template <typename Vec, typename T>
using if_elem = typename std::enable_if<std::is_same<typename elem_of<Vec>::type, T>::value, int>::type;

// one context per device for the whole process (the reference has no context object)
inline grace_b200_ctx* context()
{
    static grace_b200_ctx* ctxs[64] = {};
    int dev = 0;
    GRACE_CUDA_CHECK(cudaGetDevice(&dev));
    if (!ctxs[dev]) GRACE_B200_CHECK(grace_b200_create(&ctxs[dev], dev));
    return ctxs[dev];
}
} // namespace detail

} // namespace grace
