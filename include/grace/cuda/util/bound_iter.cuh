// grace/cuda/util/bound_iter.cuh -- the iterator user functors receive for their shared-memory
// block (reference: cuda/util/bound_iter.cuh:14-230): a typed pointer that remembers the extent
// of the allocation (checked only when GRACE_DEBUG is defined) and converts between element types.
#pragma once
#include "grace/types.h"

#include <cstddef>
#include <iterator>

namespace grace {
namespace gpu {

template <typename T>
class BoundIter {
    char* alloc_end;
    T* ptr;
    template <typename U> friend class BoundIter;

public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef T value_type;
    typedef ptrdiff_t difference_type;
    typedef T* pointer;
    typedef T& reference;

    __device__ BoundIter(char* const begin, const size_t bytes) : alloc_end(begin + bytes), ptr(reinterpret_cast<T*>(begin)) {}
    template <typename U>
    __device__ BoundIter(const BoundIter<U>& other) : alloc_end(other.alloc_end), ptr(reinterpret_cast<T*>(other.ptr)) {}
    template <typename U>
    __device__ BoundIter<T>& operator=(const BoundIter<U>& other)
    {
        alloc_end = other.alloc_end;
        ptr = reinterpret_cast<T*>(other.ptr);
        return *this;
    }
    __device__ T& operator*() const { return *ptr; }
    __device__ T& operator[](difference_type i) const { return ptr[i]; }
    __device__ T* operator->() const { return ptr; }
    __device__ BoundIter<T>& operator++() { ++ptr; return *this; }
    __device__ BoundIter<T> operator++(int) { BoundIter<T> t = *this; ++ptr; return t; }
    __device__ BoundIter<T>& operator--() { --ptr; return *this; }
    __device__ BoundIter<T> operator--(int) { BoundIter<T> t = *this; --ptr; return t; }
    __device__ BoundIter<T>& operator+=(difference_type n) { ptr += n; return *this; }
    __device__ BoundIter<T>& operator-=(difference_type n) { ptr -= n; return *this; }
    __device__ BoundIter<T> operator+(difference_type n) const { BoundIter<T> t = *this; t.ptr += n; return t; }
    __device__ BoundIter<T> operator-(difference_type n) const { BoundIter<T> t = *this; t.ptr -= n; return t; }
    __device__ difference_type operator-(const BoundIter<T>& o) const { return ptr - o.ptr; }
    __device__ bool operator==(const BoundIter<T>& o) const { return ptr == o.ptr; }
    __device__ bool operator!=(const BoundIter<T>& o) const { return ptr != o.ptr; }
    __device__ bool operator<(const BoundIter<T>& o) const { return ptr < o.ptr; }
};

} // namespace gpu
} // namespace grace
