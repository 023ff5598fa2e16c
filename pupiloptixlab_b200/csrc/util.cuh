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

// Device memory pool (pb2_api.cu).  cudaMalloc / cudaFree of the multi-GB path-state and BVH arrays cost hundreds of
// milliseconds per scene reload on a B200 box (measured: 0.75-0.83 s per reload of the Cornell box, 300 ms on the first
// builds of the 30 M-triangle scene), so freed blocks are kept per device and handed out again to requests of a
// similar size.  pool_free keeps cudaFree's ordering guarantee (the device is idle when a block changes owner).
void *pool_alloc(size_t bytes);
void pool_free(void *ptr) noexcept;
void pool_trim() noexcept; // returns every cached block to the driver (pb2_trim, pb2_shutdown, allocation failure)

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
        if (count) ptr = static_cast<T *>(pool_alloc(count * sizeof(T)));
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void release() {
        if (ptr) pool_free(ptr);
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
