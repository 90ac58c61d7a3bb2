// common.cuh — shared helpers of the sm_100a ANLS library (error handling, RAII device buffers, warp helpers).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/nnlm_b200.h"

namespace nnlm {

constexpr double TINY_NUM = 1e-16;   // reference src/nnlm.h:17

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

#define NNLM_CUDA_CHECK(expr)                                                                           \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            std::snprintf(b_, sizeof b_, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e_),       \
                          cudaGetErrorString(e_), __FILE__, __LINE__, #expr);                           \
            throw ::nnlm::Error(e_ == cudaErrorMemoryAllocation ? NNLM_E_NOMEM : NNLM_E_CUDA, b_);      \
        }                                                                                               \
    } while (0)

// every kernel launch of this library is followed by NNLM_LAUNCHED(): error check + launch accounting (nnlm_stats.launches)
inline std::atomic<uint64_t>& launch_counter() { static std::atomic<uint64_t> c{0}; return c; }
#define NNLM_LAUNCHED()                                                                                 \
    do {                                                                                                \
        NNLM_CUDA_CHECK(cudaGetLastError());                                                            \
        ::nnlm::launch_counter().fetch_add(1, std::memory_order_relaxed);                               \
    } while (0)

#define NNLM_REQUIRE(cond, msg)                                                                         \
    do {                                                                                                \
        if (!(cond)) throw ::nnlm::Error(NNLM_E_ARG, std::string(msg));                                 \
    } while (0)

// Device allocations come from the device's default stream-ordered memory pool with the release threshold lifted, so the
// multi-GB buffers of one nnmf() call are recycled by the next one instead of going back to the driver (measured: the
// cudaMalloc/cudaFree round trips of a 50000 x 10000 problem cost several hundred milliseconds per call, more than the
// 20 iterations they bracket). Callers synchronise their stream before a buffer goes out of scope (Engine::sync).
inline void pool_setup_once()
{
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
}

// host wall clock spent inside cudaMallocAsync (+ its synchronisation) on this thread: nnlm_stats.host_alloc_ms
inline double& alloc_ms_counter() { static thread_local double ms = 0.0; return ms; }
inline double& free_ms_counter() { static thread_local double ms = 0.0; return ms; }          // cudaFreeAsync
inline double& pinned_ms_counter() { static thread_local double ms = 0.0; return ms; }        // cudaMallocHost / cudaFreeHost
struct ScopedMs {
    double& acc; std::chrono::steady_clock::time_point t0;
    explicit ScopedMs(double& a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~ScopedMs() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// RAII device allocation
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t count = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n) { alloc(n); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), count(o.count) { o.p = nullptr; o.count = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; count = o.count; o.p = nullptr; o.count = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        if (n == 0) return;
        pool_setup_once();
        const auto t0 = std::chrono::steady_clock::now();
        NNLM_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), (cudaStream_t)0));
        NNLM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)0));      // usable from any stream from here on
        alloc_ms_counter() += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        count = n;
    }
    void ensure(size_t n) { if (n > count) alloc(n); }
    void release() { if (p) { ScopedMs t(free_ms_counter()); cudaFreeAsync(p, (cudaStream_t)0); p = nullptr; count = 0; } }
    size_t bytes() const { return count * sizeof(T); }
    explicit operator bool() const { return p != nullptr; }
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t count = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void alloc(size_t n) {
        release();
        if (n == 0) return;
        ScopedMs t(pinned_ms_counter());
        NNLM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T)));
        count = n;
    }
    void ensure(size_t n) { if (n > count) alloc(n); }
    void release() { if (p) { ScopedMs t(pinned_ms_counter()); cudaFreeHost(p); p = nullptr; count = 0; } }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// xor-butterfly sum: every lane ends with the bit-identical total (same association tree on all lanes)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
// `!isfinite` exactly as std::isfinite / arma::is_finite: exponent field all ones
__device__ __forceinline__ bool is_missing(double a) {
    return ((__double2hiint(a) >> 20) & 0x7ff) == 0x7ff;
}
#endif

}  // namespace nnlm
