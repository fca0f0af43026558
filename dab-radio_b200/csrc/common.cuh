// Shared host/device helpers for libdab_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "dab_b200.h"

namespace dabb200 {

// thread-local error text behind dab_last_error()
std::string& last_error_ref();
int set_error(int status, const char* fmt, ...);

#define DAB_CUDA_CHECK(expr)                                                                                  \
    do {                                                                                                      \
        cudaError_t err__ = (expr);                                                                           \
        if (err__ != cudaSuccess) {                                                                           \
            return ::dabb200::set_error(DAB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                                        __FILE__, __LINE__);                                                  \
        }                                                                                                     \
    } while (0)

// Selects the device and refuses anything that is not Blackwell sm_100: there is no fallback path.
int select_device(int device);

// Restores the caller's current CUDA device when a library call returns (a single-process multi-GPU host such as PyTorch must
// not find its current device switched by a call into this library).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <typename T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    // grows only; contents are not preserved
    cudaError_t reserve(size_t n) {
        if (n <= count) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
};

template <typename T>
struct PinnedBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    ~PinnedBuffer() { release(); }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        count = 0;
    }
    cudaError_t reserve(size_t n) {
        if (n <= count) return cudaSuccess;
        release();
        cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&ptr), n * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
};

}  // namespace dabb200
