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

// ---- programmatic dependent launch (the kernels of the training step form one dependent chain) -------------------------------
// A kernel launched through launch_pdl may START while its predecessor on the stream is still finishing: its blocks become resident
// as the predecessor's blocks retire, run their prologue (index math, shared-memory carve-up, barrier init, TMEM allocation, loads of
// operands nobody in the chain writes) and park at pdl_wait() until the predecessor has completed and its writes are visible.  The
// predecessor says when that may begin with pdl_trigger() (every block: at its start -- the dependent launch fires once all blocks of
// the predecessor are resident or done, so nothing of the predecessor ever queues behind a parked block).  In a CUDA graph the launch
// becomes a programmatic dependency edge.  Without the attribute (plain <<<>>> launches, LNRF_PDL unset) both instructions are no-ops.
// MEASURED AND LEFT OFF (LNRF_PDL=1 turns it on): with the wait at the top of every kernel of the step the graph replays at 0.3996
// against 0.3978 ms without -- inside a graph the hand-over between two kernels is already ~1.5 us -- and with the software-pipelined
// step it LOSES 24 us (0.392 vs 0.368 ms): the parked blocks of the next kernel take the SM slots that the retiring blocks free, which
// is exactly where the look-ahead march of the next batch was running.
bool pdl_enabled();  // lib.cu: LNRF_PDL == "1"
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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
