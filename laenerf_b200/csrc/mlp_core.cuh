// mlp_core.cuh -- tcgen05 / TMEM / mbarrier / cp.async PTX wrappers and SWIZZLE_128B tile helpers shared by the
// fused-MLP kernels (ffmlp.cu) and the fused NeRF network kernels (nerfnet.cu).  sm_100a only.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encode entry point is fetched through the runtime, no libcuda link)
#include "common.cuh"

namespace lnrf {

constexpr uint32_t kRows = 128;                      // rows per tile == UMMA M
constexpr uint32_t kTileBytes = kRows * 128;         // one operand tile: 128 rows x 128 B
constexpr uint32_t kWBytes = 64 * 128;               // one 64-row weight tile
constexpr uint32_t kMaxLayers = 6;

struct MlpShape {
    uint32_t in_dim, out_dim, n_layers, act, out_act;
};

// glue of NeRFNetwork.forward's backward fused into the colour-net backward kernel (see k_ffmlp_bwd<ACT, GLUE = true>)
struct BwdGlue {
    const float* grad_rgb;    // [B,3] dL/drgb (fp32, as composite_rays_train's backward writes it)
    const float* rgb;         // [B,3] the saved sigmoid outputs (fp16 values held in fp32)
    const float* grad_sigma;  // [B]   dL/dsigma (fp32)
    const __half* h0;         // [B]   the sigma net's raw output channel 0 (log-density)
    float density_scale;
    __half* dh;               // [B,16] out: dL/dh of the sigma net
};

// host-side launchers shared between ffmlp.cu and nerfnet.cu
int mlp_shape(const char* who, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t num_layers, MlpShape* sh);
// partial weight-gradient sums of one network waiting for their fixed-order reduction (k_wgrad_reduce)
struct WgradPending {
    const float* partial;  // [nslices][n] fp32
    uint32_t nslices;
    __half* gw;            // [n] fp16 gradient (written, or added to when accumulate != 0)
    uint32_t n;
    int accumulate;
};
// defer != nullptr: the reduction is not launched; *defer describes it for wgrad_reduce_pair
int ffmlp_bwd_run(const char* who, const void* grad_f16, const void* inputs_f16, const void* weights_f16, const void* forward_buffer_f16,
                  uint32_t B, const MlpShape& sh, int calc_grad_inputs, void* grad_inputs_f16, void* grad_weights_f16,
                  void* wgrad_scratch, size_t wgrad_scratch_bytes, const BwdGlue* glue, int accumulate, cudaStream_t st,
                  WgradPending* defer = nullptr);
int wgrad_reduce_pair(const char* who, const WgradPending& a, const WgradPending& b, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// byte offset of 16-byte chunk c (0..7) of row r inside a SWIZZLE_128B tile (tile base 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).  8-row groups are 1024 B apart in every tile.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D=f32 [4,6), A/B=f16 (0), a_major [15], b_major [16] (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ACT template parameter of the kernels: 0 = ReLU (every LAENeRF net), kActRuntime = decided per launch from MlpShape.
// Keeping the 7-way switch out of the 64-element epilogue loops matters: with a runtime switch the epilogue, not
// the tensor core, bounds the kernel (ncu source page, profiles/r1_ffmlp_fwd_stalls.txt).
constexpr int kActRuntime = -1;

__device__ __forceinline__ float act_fwd(uint32_t a, float x) {  // ffmlp/src/utils.h:424-475
    switch (a) {
        case 0: return fmaxf(x, 0.0f);
        case 1: return __expf(x);
        case 2: return __sinf(x);
        case 3: return 1.0f / (1.0f + __expf(-x));
        case 4: { const float y = x * 10.0f; return 0.5f * (y + sqrtf(y * y + 4.0f)) / 10.0f; }
        case 5: return __logf(__expf(x * 10.0f) + 1.0f) / 10.0f;
        default: return x;
    }
}
__device__ __forceinline__ float act_bwd(uint32_t a, float g, float fwd) {  // utils.h:538-583 (through the stored output)
    switch (a) {
        case 0: return fwd > 0.0f ? g : 0.0f;
        case 1: return g * fwd;
        case 3: return g * (fwd * (1.0f - fwd));
        case 4: { const float y = fwd * 10.0f; return g * (y * y / (y * y + 1.0f)); }
        case 5: return g * (1.0f - __expf(-fwd * 10.0f));
        default: return g;
    }
}

// ---- asynchronous global -> shared copies (LDGSTS): all 16-byte pieces of a tile are in flight together ----
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows x K fp16 row-major (global) -> SWIZZLE_128B tile at shared address `tile`
__device__ __forceinline__ void load_rows_async(uint32_t tile, const __half* __restrict__ src, uint32_t rows, uint32_t K, int tid) {
    const uint32_t cpr = K >> 3, n = rows * cpr;
    const uint4* g = reinterpret_cast<const uint4*>(src);
    for (uint32_t c = tid; c < n; c += 128) {
        const uint32_t r = c / cpr, cc = c - r * cpr;
        cp_async16(tile + sw128(r, cc), g + c);
    }
}
// raw copy of n16 16-byte pieces (staging area, no swizzle)
__device__ __forceinline__ void copy_raw_async(uint32_t dst, const void* __restrict__ src, uint32_t n16, int tid) {
    const uint4* g = reinterpret_cast<const uint4*>(src);
    for (uint32_t c = tid; c < n16; c += 128) cp_async16(dst + c * 16u, g + c);
}
// shared (row-major [rows][K] fp16 at `src`) -> SWIZZLE_128B tile holding the TRANSPOSE: element (r, k) -> tile row k, column r
__device__ __forceinline__ void transpose_to_tile(uint8_t* tile, const __half* src, uint32_t rows, uint32_t K, int tid) {
    for (uint32_t e = tid; e < rows * K; e += 128) {
        const uint32_t r = e / K, k = e - r * K;
        *reinterpret_cast<__half*>(tile + sw128(k, r >> 3) + (r & 7u) * 2u) = src[e];
    }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    union { uint4 u; __half2 h[4]; } p;
#pragma unroll
    for (int j = 0; j < 4; j++) p.h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    return p.u;
}
__device__ __forceinline__ void unpack8(uint4 u, float* v) {
    union { uint4 u; __half2 h[4]; } p;
    p.u = u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float2 f = __half22float2(p.h[j]);
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
}

// 32 consecutive fp32 columns of this thread's TMEM lane, WITHOUT waiting (pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- helpers of the multi-tile-in-flight kernels (one warpgroup of 128 threads per tile, several tiles per CTA) ----

__device__ __forceinline__ void group_barrier(uint32_t g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1u) : "memory"); }

// mbarrier wait for the hot loop: try_wait suspends in hardware; the wall-clock bound is only consulted every 64 K polls
__device__ __forceinline__ void mbar_wait_hot(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t polls = 0;
    long long t0 = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if ((++polls & 0xffffu) == 0u) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000ll) __trap();
        }
    }
}

// two fp32 -> packed fp16x2 with ReLU in one instruction (F2FP.RELU); `lo` lands in the low half (lower address)
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// one full 32-byte sector per thread, streaming (evict-first)
__device__ __forceinline__ void st_global_32B(void* p, const uint32_t* r) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// rows x K fp16 row-major (global) -> SWIZZLE_128B tile, issued by the `nthr` threads of one group / of the CTA
__device__ __forceinline__ void load_rows_async_n(uint32_t tile, const __half* __restrict__ src, uint32_t rows, uint32_t K, uint32_t t,
                                                  uint32_t nthr) {
    const uint32_t cpr = K >> 3, n = rows * cpr;
    const uint4* g = reinterpret_cast<const uint4*>(src);
    for (uint32_t c = t; c < n; c += nthr) {
        const uint32_t r = c / cpr, cc = c - r * cpr;
        cp_async16(tile + sw128(r, cc), g + c);
    }
}

// ---- MMA issue: warp-uniform control flow + elect.sync + unrolled K loop ----
// Guarding the issue with `tid == 0` makes ptxas wrap EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (the
// operands must be uniform registers and the branch is divergent to it): ~30 instructions and ~150-200 cycles per MMA, which
// put 0.4 us (forward, 4 MMAs) to 1.2 us (backward, 12 MMAs) of pure issue overhead on every chain step (SASS of profiles/r1d).
// With the whole warp taking the branch and elect.sync as the inner predicate the MMAs are emitted back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// NK MMAs whose operand descriptors advance by (step_a, step_b) encoded units per K-slice; the first accumulates iff acc_first
template <int NK>
__device__ __forceinline__ void umma_chain(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t step_a, uint32_t step_b, uint32_t idesc,
                                           bool acc_first) {
#pragma unroll
    for (int k = 0; k < NK; k++) umma_f16(d_tmem, a + (uint64_t)(step_a * k), b + (uint64_t)(step_b * k), idesc, (k > 0 || acc_first) ? 1u : 0u);
}

// runtime K-slice count 1..4 (K = 16..64), K-major operands
__device__ __forceinline__ void umma_chain_k(uint32_t nk, uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc) {
    switch (nk) {
        case 1: umma_chain<1>(d_tmem, a, b, 2, 2, idesc, false); break;
        case 2: umma_chain<2>(d_tmem, a, b, 2, 2, idesc, false); break;
        case 3: umma_chain<3>(d_tmem, a, b, 2, 2, idesc, false); break;
        default: umma_chain<4>(d_tmem, a, b, 2, 2, idesc, false); break;
    }
}

// ---- TMA (bulk tensor copies): one elected thread moves a whole 128-row x 128-byte SWIZZLE_128B tile ----
// [rows, 64] fp16 row-major global tensor, box = one tile; defined in ffmlp.cu
int make_tile_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, const char* who);
// generic 2-D fp16 map, SWIZZLE_128B, zero fill outside the tensor; defined in ffmlp.cu
int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_bytes, uint32_t box_cols,
                       uint32_t box_rows, const char* who);

__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, uint32_t smem_tile, int32_t col, int32_t row) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_tile), "r"(col), "r"(row)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace lnrf
