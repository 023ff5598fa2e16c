// Host-side helpers shared by the pb2 translation units: error capture, RAII device buffers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace pb2 {
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        throw CudaError(buf);
    }
}
#define PB2_CUDA(x) ::pb2::cuda_check((x), #x, __FILE__, __LINE__)
#define PB2_LAUNCH_CHECK() ::pb2::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

template<typename T>
struct DevBuf {
    T *ptr = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : ptr(o.ptr), n(o.n) { o.ptr = nullptr, o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            ptr = o.ptr, n = o.n, o.ptr = nullptr, o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) PB2_CUDA(cudaMalloc(reinterpret_cast<void **>(&ptr), count * sizeof(T)));
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
    void upload(const T *host, size_t count, cudaStream_t s = 0) {
        ensure(count);
        if (count) PB2_CUDA(cudaMemcpyAsync(ptr, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(cudaStream_t s = 0) {
        if (n) PB2_CUDA(cudaMemsetAsync(ptr, 0, bytes(), s));
    }
};

inline unsigned div_up(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }
}// namespace pb2
