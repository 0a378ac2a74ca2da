// ffmlp.cu -- fully fused 64-wide MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces ffmlp/src/ffmlp.cu of the reference: kernel_mlp_fused (:331-407, wmma fp16-accumulate),
// kernel_mlp_fused_backward (:410-518) and the CUTLASS 2.8 split-K weight-gradient GEMMs (:783-887).
//
// Design:
//   * a CTA of 128 threads owns tiles of 128 batch rows = the native tcgen05.mma M.  Every operand tile lives in
//     shared memory as 128-byte rows (64 fp16) in the SWIZZLE_128B canonical layout, so the SAME tile can be read
//     as a K-major operand (forward / dgrad: K = feature axis) and as an MN-major operand (wgrad: K = batch rows).
//   * all weight matrices stay resident in shared memory for the life of the CTA (forward: W_m as [N, K] K-major;
//     backward: W_m^T), accumulators live in TMEM (fp32), one elected thread issues the MMAs and signals an
//     mbarrier through tcgen05.commit, the 128 threads read their accumulator row back with tcgen05.ld
//     (thread r <-> TMEM lane r), apply the activation and write the fp16 row of the next layer's A operand.
//   * backward: dL/dhidden is chained on chip the same way (ReLU mask from the saved activations), and every
//     layer's weight gradient is ONE TMEM accumulator that is summed over all tiles the CTA processes
//     (dW = sum over tiles of X_tile^T . G_tile, M = feature, N = feature, K = 128 rows), flushed once per CTA
//     with coalesced fp32 reductions into a scratch vector -- no split-K workspace, no side streams.
//   Accumulation is fp32 everywhere (the reference accumulates in fp16); activations are rounded to fp16 exactly
//   where the reference stores them (forward_buffer, backward chain, outputs).
#include "common.cuh"

namespace lnrf {

constexpr uint32_t kRows = 128;                      // rows per tile == UMMA M
constexpr uint32_t kTileBytes = kRows * 128;         // one operand tile: 128 rows x 128 B
constexpr uint32_t kWBytes = 64 * 128;               // one 64-row weight tile
constexpr uint32_t kMaxLayers = 6;

struct MlpShape {
    uint32_t in_dim, out_dim, n_layers, act, out_act;
};

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

// rows x K fp16 row-major (global) -> SWIZZLE_128B tile; TRANSPOSE stores element (r, k) at tile row k, column r
__device__ __forceinline__ void load_rows(uint8_t* tile, const __half* __restrict__ src, uint32_t rows, uint32_t K, int tid) {
    const uint32_t cpr = K >> 3;
    for (uint32_t c = tid; c < rows * cpr; c += 128) {
        const uint32_t r = c / cpr, cc = c - r * cpr;
        *reinterpret_cast<uint4*>(tile + sw128(r, cc)) = __ldg(reinterpret_cast<const uint4*>(src) + c);
    }
}
__device__ __forceinline__ void load_rows_transposed(uint8_t* tile, const __half* __restrict__ src, uint32_t rows, uint32_t K, int tid) {
    for (uint32_t e = tid; e < rows * K; e += 128) {
        const uint32_t r = e / K, k = e - r * K;  // src[r][k] -> tile row k, column r
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

// =========================================================================================================
// forward / inference
// =========================================================================================================
template <bool TRAIN>
__global__ void __launch_bounds__(128)
k_ffmlp_fwd(const __half* __restrict__ inputs, const __half* __restrict__ weights, __half* __restrict__ fwd_buf,
            __half* __restrict__ outputs, const uint32_t B, const MlpShape sh, const uint32_t ntiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    const uint32_t NL = sh.n_layers, in_dim = sh.in_dim, out_dim = sh.out_dim;
    uint8_t* sW = sm;                                        // W_0 .. W_{NL-1}: 8 KB each, W_NL: out_dim rows
    uint8_t* sA = sW + NL * kWBytes + ((out_dim * 128u + 1023u) & ~1023u);  // two activation tiles (ping-pong)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sA + 2 * kTileBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t row = tid;  // this thread's row of the tile == its TMEM lane

    if (warp == 0) tmem_alloc(tslot, 64);
    if (tid == 32) { mbar_init(mbar, 1); fence_mbar_init(); }
    load_rows(sW, weights, 64, in_dim, tid);
    for (uint32_t m = 1; m < NL; m++) load_rows(sW + m * kWBytes, weights + 64 * in_dim + (m - 1) * 4096, 64, 64, tid);
    load_rows(sW + NL * kWBytes, weights + 64 * in_dim + (NL - 1) * 4096, out_dim, 64, tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t r0 = (size_t)tile * kRows;
        load_rows(sA, inputs + r0 * in_dim, kRows, in_dim, tid);
        fence_proxy_async();
        __syncthreads();
        for (uint32_t m = 0; m <= NL; m++) {
            uint8_t* cur = sA + (m & 1u) * kTileBytes;
            uint8_t* nxt = sA + ((m + 1u) & 1u) * kTileBytes;
            if (tid == 0) {
                tc_fence_after();
                const uint32_t K = m == 0 ? in_dim : 64u, N = m == NL ? out_dim : 64u;
                const uint32_t idesc = make_idesc(128, N, false, false);
                const uint64_t a = desc_sw128(smem_u32(cur), 16), b = desc_sw128(smem_u32(sW + m * kWBytes), 16);
                for (uint32_t k = 0; k < K / 16; k++) umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, k > 0);
                umma_commit(mbar);
            }
            if (TRAIN && m > 0) {  // save H_{m-1} (== cur) while the tensor core works: coalesced 16-byte rows
                uint4* dst = reinterpret_cast<uint4*>(fwd_buf + ((size_t)(m - 1) * B + r0) * 64);
#pragma unroll
                for (uint32_t i = 0; i < 8; i++) {
                    const uint32_t c = tid + i * 128;
                    __stcs(dst + c, *reinterpret_cast<const uint4*>(cur + sw128(c >> 3, c & 7u)));
                }
            }
            mbar_wait(mbar, phase);
            phase ^= 1u;
            tc_fence_after();
            if (m < NL) {
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
                    float v[16];
                    tmem_ld16(taddr + q * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = act_fwd(sh.act, v[i]);
                    *reinterpret_cast<uint4*>(nxt + sw128(row, 2 * q)) = pack8(v);
                    *reinterpret_cast<uint4*>(nxt + sw128(row, 2 * q + 1)) = pack8(v + 8);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncthreads();
            } else {
                __half* o = outputs + (r0 + row) * out_dim;
                for (uint32_t q = 0; q < out_dim / 16; q++) {
                    float v[16];
                    tmem_ld16(taddr + q * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = act_fwd(sh.out_act, v[i]);
                    __stcs(reinterpret_cast<uint4*>(o + q * 16), pack8(v));
                    __stcs(reinterpret_cast<uint4*>(o + q * 16) + 1, pack8(v + 8));
                }
                tc_fence_before();
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// =========================================================================================================
// backward: dL/dhidden chain + all weight gradients + optional dL/dinput, one kernel
// =========================================================================================================
__device__ __host__ inline uint32_t bwd_tmem_cols(uint32_t in_dim, uint32_t out_dim, uint32_t NL) {
    const uint32_t need = 64 + out_dim + 64 * (NL - 1) + in_dim;
    uint32_t c = 32;
    while (c < need) c <<= 1;
    return c;
}

__global__ void __launch_bounds__(128)
k_ffmlp_bwd(const __half* __restrict__ grad, const __half* __restrict__ inputs, const __half* __restrict__ weights,
            const __half* __restrict__ fwd_buf, __half* __restrict__ grad_inputs, float* __restrict__ wgrad, const uint32_t B,
            const MlpShape sh, const uint32_t ntiles, const int calc_grad_inputs) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    const uint32_t NL = sh.n_layers, in_dim = sh.in_dim, out_dim = sh.out_dim;
    // W_m^T tiles: index m-1 for m = 1..NL ([64 rows = input feature][K = n_m]); then W_0^T ([in rows][K = 64])
    uint8_t* sWT = sm;
    uint8_t* sWT0 = sWT + NL * kWBytes;
    uint8_t* sX = sWT0 + kWBytes;              // order matters: every tile used as an M=128 MN-major A operand
    uint8_t* sH = sX + kTileBytes;             // is followed by another tile (the ignored second atom, rows 64..127 of D)
    uint8_t* sG = sH + NL * kTileBytes;        // G_m lives in sG[m & 1]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sG + 2 * kTileBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t row = tid;
    const uint32_t ncols = bwd_tmem_cols(in_dim, out_dim, NL);

    if (warp == 0) tmem_alloc(tslot, ncols);
    if (tid == 32) { mbar_init(mbar, 1); fence_mbar_init(); }
    const __half* W_last = weights + 64 * in_dim + (NL - 1) * 4096;
    load_rows_transposed(sWT + (NL - 1) * kWBytes, W_last, out_dim, 64, tid);  // W_NL [out,64] -> [64][out]
    for (uint32_t m = 1; m < NL; m++) load_rows_transposed(sWT + (m - 1) * kWBytes, weights + 64 * in_dim + (m - 1) * 4096, 64, 64, tid);
    if (calc_grad_inputs) load_rows_transposed(sWT0, weights, 64, in_dim, tid);  // W_0 [64,in] -> [in][64]
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    // TMEM columns: [0,64) dgrad accumulator; then dW_NL (out_dim), dW_{NL-1}..dW_1 (64 each), dW_0 (in_dim)
    auto acc_col = [&](uint32_t m) -> uint32_t {
        if (m == NL) return 64u;
        if (m == 0) return 64u + out_dim + 64u * (NL - 1);
        return 64u + out_dim + 64u * (NL - 1 - m);
    };
    uint32_t phase = 0, iter = 0;

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, iter++) {
        const size_t r0 = (size_t)tile * kRows;
        load_rows(sG + (NL & 1u) * kTileBytes, grad + r0 * out_dim, kRows, out_dim, tid);
        load_rows(sX, inputs + r0 * in_dim, kRows, in_dim, tid);
        for (uint32_t l = 0; l < NL; l++) load_rows(sH + l * kTileBytes, fwd_buf + ((size_t)l * B + r0) * 64, kRows, 64, tid);
        fence_proxy_async();
        __syncthreads();
        for (uint32_t m = NL; m >= 1; m--) {
            uint8_t* Gm = sG + (m & 1u) * kTileBytes;
            uint8_t* Gp = sG + ((m - 1u) & 1u) * kTileBytes;
            uint8_t* Hp = sH + (m - 1) * kTileBytes;  // H_{m-1}: activation mask AND the input of matmul m
            const uint32_t n_m = m == NL ? out_dim : 64u;
            if (tid == 0) {
                tc_fence_after();
                // dgrad: D[128,64] = G_m[128,n_m] . W_m[n_m,64]
                const uint64_t a = desc_sw128(smem_u32(Gm), 16), b = desc_sw128(smem_u32(sWT + (m - 1) * kWBytes), 16);
                const uint32_t idesc = make_idesc(128, 64, false, false);
                for (uint32_t k = 0; k < n_m / 16; k++) umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, k > 0);
                umma_commit(mbar);
                // wgrad: acc_m[feature i][neuron j] += H_{m-1}^T . G_m  (K = the 128 rows, 16 per MMA = 2048 B)
                const uint64_t wa = desc_sw128(smem_u32(Hp), kTileBytes), wb = desc_sw128(smem_u32(Gm), kTileBytes);
                const uint32_t widesc = make_idesc(128, n_m, true, true);
                for (uint32_t k = 0; k < 8; k++) umma_f16(tmem + acc_col(m), wa + 128 * k, wb + 128 * k, widesc, (iter | k) > 0);
            }
            mbar_wait(mbar, phase);
            phase ^= 1u;
            tc_fence_after();
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) {
                float v[16], h[16];
                tmem_ld16(taddr + q * 16, v);
                unpack8(*reinterpret_cast<const uint4*>(Hp + sw128(row, 2 * q)), h);
                unpack8(*reinterpret_cast<const uint4*>(Hp + sw128(row, 2 * q + 1)), h + 8);
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = act_bwd(sh.act, v[i], h[i]);
                *reinterpret_cast<uint4*>(Gp + sw128(row, 2 * q)) = pack8(v);
                *reinterpret_cast<uint4*>(Gp + sw128(row, 2 * q + 1)) = pack8(v + 8);
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
        }
        // input layer: dW_0[neuron j][feature i] += G_0^T . X ; optionally dX = G_0 . W_0
        if (tid == 0) {
            tc_fence_after();
            if (calc_grad_inputs) {
                const uint64_t a = desc_sw128(smem_u32(sG), 16), b = desc_sw128(smem_u32(sWT0), 16);
                const uint32_t idesc = make_idesc(128, in_dim, false, false);
                for (uint32_t k = 0; k < 4; k++) umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, k > 0);
            }
            const uint64_t wa = desc_sw128(smem_u32(sG), kTileBytes), wb = desc_sw128(smem_u32(sX), kTileBytes);
            const uint32_t widesc = make_idesc(128, in_dim, true, true);
            for (uint32_t k = 0; k < 8; k++) umma_f16(tmem + acc_col(0), wa + 128 * k, wb + 128 * k, widesc, (iter | k) > 0);
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase);  // every MMA of this tile has finished: shared-memory tiles may be overwritten
        phase ^= 1u;
        tc_fence_after();
        if (calc_grad_inputs) {
            __half* gi = grad_inputs + (r0 + row) * in_dim;
            for (uint32_t q = 0; q < in_dim / 16; q++) {
                float v[16];
                tmem_ld16(taddr + q * 16, v);
                __stcs(reinterpret_cast<uint4*>(gi + q * 16), pack8(v));
                __stcs(reinterpret_cast<uint4*>(gi + q * 16) + 1, pack8(v + 8));
            }
        }
        tc_fence_before();
        __syncthreads();
    }

    // ---- flush the weight-gradient accumulators (rows 0..63 are real; lanes 64..127 hold the ignored atom) ----
    tc_fence_after();
    if (iter > 0 && warp < 2) {
        const uint32_t w_first = 64 * in_dim;
        for (uint32_t m = 1; m <= NL; m++) {  // acc_m[i][j] = dW_m[j][i]; lane = i -> coalesced over i
            const uint32_t n_m = m == NL ? out_dim : 64u;
            float* dst = wgrad + w_first + (m - 1) * 4096;
            for (uint32_t q = 0; q < n_m / 16; q++) {
                float v[16];
                tmem_ld16(taddr + acc_col(m) + q * 16, v);
#pragma unroll
                for (int j = 0; j < 16; j++) atomicAdd(dst + (q * 16 + j) * 64 + row, v[j]);
            }
        }
        for (uint32_t q = 0; q < in_dim / 16; q++) {  // acc_0[j][i] = dW_0[j][i]; lane = j
            float v[16];
            tmem_ld16(taddr + acc_col(0) + q * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i++) atomicAdd(wgrad + row * in_dim + q * 16 + i, v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

__global__ void __launch_bounds__(256) k_ffmlp_wgrad_finalize(const float* __restrict__ acc, __half* __restrict__ gw, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gw[i] = __float2half_rn(acc[i]);
}

}  // namespace lnrf

using namespace lnrf;

static int check_mlp(const char* who, uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                     uint32_t activation, uint32_t output_activation, MlpShape* sh) {
    if (!(hidden_dim == 16 || hidden_dim == 32 || hidden_dim == 64 || hidden_dim == 128 || hidden_dim == 256)) {
        set_error("hidden_dim should in [16, 32, 64, 128, 256]");  // ffmlp.cu:657
        return LNRF_ERR_INVALID_ARGUMENT;
    }
    if (hidden_dim != 64) {
        set_error("%s: hidden_dim %u not built (this library implements the 64-wide nets LAENeRF uses)", who, hidden_dim);
        return LNRF_ERR_UNSUPPORTED;
    }
    LNRF_REQUIRE(input_dim > 0 && input_dim % 16 == 0, "FFMLP input_dim should be 16 * m (m > 0), but got %u", input_dim);
    LNRF_REQUIRE(output_dim > 0 && output_dim % 16 == 0, "%s: (padded) output_dim should be 16 * m, but got %u", who, output_dim);
    LNRF_REQUIRE(num_layers >= 2, "FFMLP num_layers should be larger than 2 (3 matmuls), but got %u", num_layers);
    LNRF_REQUIRE(B % 128 == 0, "ffmlp batch size must be 128 * m (m > 0), but got %u.", B);
    if (input_dim > 64 || output_dim > 64 || num_layers > kMaxLayers) {
        set_error("%s: input_dim %u / output_dim %u / num_layers %u outside the built range (<=64, <=64, <=%u)", who, input_dim,
                  output_dim, num_layers, kMaxLayers);
        return LNRF_ERR_UNSUPPORTED;
    }
    LNRF_REQUIRE(activation <= 6 && output_activation <= 6, "%s: activation id out of range", who);
    sh->in_dim = input_dim; sh->out_dim = output_dim; sh->n_layers = num_layers; sh->act = activation; sh->out_act = output_activation;
    return LNRF_OK;
}

static size_t fwd_smem_bytes(const MlpShape& sh) {
    return 1024 + sh.n_layers * kWBytes + ((sh.out_dim * 128u + 1023u) & ~1023u) + 2 * kTileBytes + 64;
}
static size_t bwd_smem_bytes(const MlpShape& sh) {
    return 1024 + (sh.n_layers + 1) * kWBytes + kTileBytes * (1 + sh.n_layers + 2) + 64;
}

template <bool TRAIN>
static int ffmlp_fwd_launch(const char* who, const void* inputs, const void* weights, uint32_t B, const MlpShape& sh, void* fwd_buf,
                            void* outputs, cudaStream_t st) {
    if (B == 0) return LNRF_OK;
    LNRF_REQUIRE(inputs && weights && outputs && (!TRAIN || fwd_buf), "%s: null pointer", who);
    LNRF_REQUIRE(((reinterpret_cast<uintptr_t>(inputs) | reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(outputs) |
                   reinterpret_cast<uintptr_t>(fwd_buf)) & 15) == 0, "%s: tensors must be 16-byte aligned", who);
    const size_t smem = fwd_smem_bytes(sh);
    cudaError_t e = cudaFuncSetAttribute(k_ffmlp_fwd<TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    const uint32_t ntiles = B / kRows;
    const uint32_t per_sm = (uint32_t)((227 * 1024) / (smem + 1024));
    const uint32_t cap = (uint32_t)kNumSMs * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    const uint32_t grid = ntiles < cap ? ntiles : cap;
    k_ffmlp_fwd<TRAIN><<<grid, 128, smem, st>>>((const __half*)inputs, (const __half*)weights, (__half*)fwd_buf, (__half*)outputs, B, sh, ntiles);
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

extern "C" {

int lnrf_ffmlp_forward(const void* inputs_f16, const void* weights_f16, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                       uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                       void* forward_buffer_f16, void* outputs_f16, lnrf_stream_t stream) {
    MlpShape sh;
    if (int e = check_mlp("ffmlp_forward", B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh)) return e;
    return ffmlp_fwd_launch<true>("ffmlp_forward", inputs_f16, weights_f16, B, sh, forward_buffer_f16, outputs_f16,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int lnrf_ffmlp_inference(const void* inputs_f16, const void* weights_f16, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                         uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                         void* inference_buffer_f16, void* outputs_f16, lnrf_stream_t stream) {
    (void)inference_buffer_f16;
    MlpShape sh;
    if (int e = check_mlp("ffmlp_inference", B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh)) return e;
    return ffmlp_fwd_launch<false>("ffmlp_inference", inputs_f16, weights_f16, B, sh, nullptr, outputs_f16,
                                   reinterpret_cast<cudaStream_t>(stream));
}

size_t lnrf_ffmlp_wgrad_scratch_bytes(uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers) {
    return sizeof(float) * (size_t)hidden_dim * (input_dim + (size_t)hidden_dim * (num_layers - 1) + output_dim);
}

int lnrf_ffmlp_backward(const void* grad_f16, const void* inputs_f16, const void* weights_f16, const void* forward_buffer_f16,
                        uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                        uint32_t activation, uint32_t output_activation, int calc_grad_inputs, void* backward_buffer_f16,
                        void* grad_inputs_f16, void* grad_weights_f16, void* wgrad_scratch, size_t wgrad_scratch_bytes,
                        lnrf_stream_t stream) {
    (void)backward_buffer_f16;  // dL/dhidden never leaves the chip
    MlpShape sh;
    if (int e = check_mlp("ffmlp_backward", B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh)) return e;
    LNRF_REQUIRE(grad_f16 && inputs_f16 && weights_f16 && forward_buffer_f16 && grad_weights_f16, "ffmlp_backward: null pointer");
    LNRF_REQUIRE(!calc_grad_inputs || grad_inputs_f16, "ffmlp_backward: calc_grad_inputs without grad_inputs");
    const size_t need = lnrf_ffmlp_wgrad_scratch_bytes(input_dim, output_dim, hidden_dim, num_layers);
    if (!wgrad_scratch || wgrad_scratch_bytes < need) {
        set_error("ffmlp_backward: wgrad scratch too small (%zu < %zu)", wgrad_scratch_bytes, need);
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    LNRF_REQUIRE(bwd_tmem_cols(input_dim, output_dim, num_layers) <= 512, "ffmlp_backward: network too deep for one TMEM allocation");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t nparams = (uint32_t)(need / sizeof(float));
    cudaError_t e = cudaMemsetAsync(wgrad_scratch, 0, need, st);
    if (e != cudaSuccess) return cuda_fail(e, "ffmlp_backward: memset");
    if (B > 0) {
        const size_t smem = bwd_smem_bytes(sh);
        LNRF_REQUIRE(smem <= 227 * 1024, "ffmlp_backward: network needs %zu B of shared memory (> 227 KiB)", smem);
        e = cudaFuncSetAttribute(k_ffmlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "ffmlp_backward: smem attribute");
        const uint32_t ntiles = B / kRows;
        const uint32_t per_sm_smem = (uint32_t)((227 * 1024) / (smem + 1024));
        const uint32_t per_sm_tmem = 512u / bwd_tmem_cols(input_dim, output_dim, num_layers);
        uint32_t per_sm = per_sm_smem < per_sm_tmem ? per_sm_smem : per_sm_tmem;
        if (per_sm < 1) per_sm = 1;
        const uint32_t cap = (uint32_t)kNumSMs * per_sm;
        const uint32_t grid = ntiles < cap ? ntiles : cap;
        k_ffmlp_bwd<<<grid, 128, smem, st>>>((const __half*)grad_f16, (const __half*)inputs_f16, (const __half*)weights_f16,
                                             (const __half*)forward_buffer_f16, (__half*)grad_inputs_f16, (float*)wgrad_scratch, B, sh,
                                             ntiles, calc_grad_inputs);
        LNRF_LAUNCH_CHECK("ffmlp_backward");
    }
    k_ffmlp_wgrad_finalize<<<div_up(nparams, 256u), 256, 0, st>>>((const float*)wgrad_scratch, (__half*)grad_weights_f16, nparams);
    LNRF_LAUNCH_CHECK("ffmlp_backward(finalize)");
    return LNRF_OK;
}

int lnrf_allocate_splitk(size_t size) { (void)size; return LNRF_OK; }
int lnrf_free_splitk(void) { return LNRF_OK; }

}  // extern "C"
