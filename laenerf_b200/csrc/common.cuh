// common.cuh -- shared host/device helpers for liblaenerf_b200 (sm_100a only, no torch headers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/laenerf_b200.h"

namespace lnrf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids for persistent kernels are multiples of this

// ---- error plumbing (thread-local message, TORCH_CHECK-like behaviour lives in the Python shim) ----------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launch_count;

#define LNRF_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            ::lnrf::set_error(__VA_ARGS__);      \
            return LNRF_ERR_INVALID_ARGUMENT;    \
        }                                        \
    } while (0)

// After every launch: count it and surface configuration errors immediately (the reference checks nothing).
#define LNRF_LAUNCH_CHECK(name)                                            \
    do {                                                                   \
        ::lnrf::g_launch_count.fetch_add(1, std::memory_order_relaxed);    \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return ::lnrf::cuda_fail(e__, name);       \
    } while (0)

template <typename T>
__host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

// ---- warp helpers --------------------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// streaming (evict-first) stores for write-once sample buffers
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float2* p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float4* p, float4 v) { __stcs(p, v); }

}  // namespace lnrf
