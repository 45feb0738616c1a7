// Shared host/device helpers for libtq100 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/tq100.h"

namespace tq {

constexpr float kTiny = 1e-8f;   // the reference's clamp floor (quantizer.py:66,100,125,240; gptq.py:177)

// ---- error reporting -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define TQ_CHECK_ARG(cond, ...)                          \
    do {                                                 \
        if (!(cond)) {                                   \
            ::tq::set_error(__VA_ARGS__);                \
            return TQ_E_BADARG;                          \
        }                                                \
    } while (0)

// every kernel launch goes through here: counts it and surfaces launch-configuration errors
#define TQ_LAUNCH_CHECK(name)                                                        \
    do {                                                                             \
        ::tq::g_launches.fetch_add(1, std::memory_order_relaxed);                    \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::tq::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

#define TQ_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ::tq::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: `mask` (one static per kernel or
// kernel family at the call site) remembers the devices it has been set on, so a process driving several GPUs sets it
// on each of them once.
static inline bool dyn_smem_pending(std::atomic<unsigned long long>& mask, int& dev) {
    dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= 64 || !((mask.load(std::memory_order_relaxed) >> dev) & 1ull);
}
static inline void dyn_smem_done(std::atomic<unsigned long long>& mask, int dev) {
    if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_relaxed);
}

int sm_count();

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

}  // namespace tq
