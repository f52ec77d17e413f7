// Shared device helpers for libpwr_b200 (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pwr.h"

namespace pwr {

constexpr int kMap = PWR_MAP_ELEMS;      // 64*64 pixels per map
constexpr int kLabel = PWR_LABEL_SIZE;   // 64
constexpr int kImage = PWR_IMAGE_SIZE;   // 128
constexpr int kThreads = 256;            // 8 warps: 16 pixels (4 float4) per thread per map
constexpr int kWarps = kThreads / 32;
constexpr int kVec = kMap / 4 / kThreads; // float4 chunks per thread per map = 4

// ---- argument checking ---------------------------------------------------
static inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

#define PWR_REQUIRE_PTR(p) do { if ((p) == nullptr) return PWR_E_NULL; if (pwr::misaligned(p)) return PWR_E_ALIGN; } while (0)
#define PWR_OPTIONAL_PTR(p) do { if ((p) != nullptr && pwr::misaligned(p)) return PWR_E_ALIGN; } while (0)

static inline int launch_status() { return static_cast<int>(cudaGetLastError()); }

// ---- per-device facts, looked up once ----------------------------------------
// The launch path must not pay a driver query per call (at batch 128-256 the kernels last tens of
// microseconds): the SM count is cached per device, and the dynamic-shared-memory opt-in of a kernel is
// made once per (kernel instantiation, device).  Benign race: two threads may both query / set once.
static inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }
static inline int sm_count(int dev) {
    static std::atomic<int> cache[64];
    const bool cached = dev >= 0 && dev < 64;
    if (cached) { const int v = cache[dev].load(std::memory_order_relaxed); if (v > 0) return v; }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (cached) cache[dev].store(sms, std::memory_order_relaxed);
    return sms;
}
#define PWR_ENSURE_DYN_SMEM(bytes, dev, ...)                                                               \
    do {                                                                                                   \
        static std::atomic<unsigned long long> done_{0ull};                                                \
        const unsigned long long bit_ = 1ull << ((dev) & 63);                                              \
        if (!(done_.load(std::memory_order_acquire) & bit_)) {                                             \
            cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes));       \
            done_.fetch_or(bit_, std::memory_order_release);                                               \
        }                                                                                                  \
    } while (0)

// ---- memory access with cache hints ----------------------------------------
// Streams that are touched exactly once (logits in, gradients out) use the
// evict-first policy so they do not push the per-sample label/mask planes
// (re-read by the J CTAs of a sample) out of L2.
__device__ __forceinline__ float4 ld_stream(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld_keep(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

// ---- reductions ----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of N values per thread; every thread gets the totals.
// Fixed shuffle tree + fixed warp order => bitwise deterministic.
// `scratch` holds kWarps*N floats and may be reused after the call returns
// (the trailing __syncthreads protects it).
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[warp * N + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) s += scratch[wv * N + i];
        v[i] = s;
    }
    __syncthreads();
}

}  // namespace pwr
