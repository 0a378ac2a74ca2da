// nerfnet.cu -- NeRFNetwork.forward / .backward of the reference (nerf/network_ff.py:51-79) as fused sm_100a kernels
// (row f-1 of SURVEY.md section 8: "SH direction encoder + the elementwise glue fused into the MLP kernels").
//
// The reference evaluates, per sample, between the hash-grid encoder and the compositor:
//     h = sigma_net(enc)                     FFMLP 32 -> 64 x n_s -> 16       (ffmlp/src/ffmlp.cu:331-407)
//     sigma = trunc_exp(h[..., 0])           fp32 exp                         (activation.py:5-17)
//     d = SHEncoder(dirs)                    degree 4, 16 channels            (shencoder/src/shencoder.cu:27-123)
//     rgb = sigmoid(color_net(cat[d, h[..., 1:], 0]))   FFMLP 32 -> 64 x n_c -> 16, 3 used
// as two MLP launches plus ~10 elementwise / cat / cast passes over [M, 16..32] tensors.  Here ONE kernel runs both
// nets per 128-sample tile on the tensor cores (tcgen05, TMEM accumulators, weights of both nets resident in shared
// memory), with the glue as register-level epilogues: the sigma epilogue applies exp, evaluates the SH basis of the
// sample's direction and assembles the colour net's input row in shared memory; the colour epilogue applies the
// sigmoid.  Values are rounded to fp16 exactly where the reference's tensors are fp16 (h, d, hidden activations,
// the colour net's output, rgb), so the result is the reference's, not an approximation of it.
//
// The backward is two launches of the FFMLP backward kernel (ffmlp.cu): the colour net with its glue fused
// (k_ffmlp_bwd<0, true>: sigmoid' prologue; dL/dgeo_feat + trunc_exp backward epilogue -> dL/dh), then the sigma
// net, whose dL/dinput feeds the hash-grid backward.
#include "mlp_core.cuh"
#include "sh_core.cuh"

namespace lnrf {

constexpr uint32_t kEncDim = 32;   // hash-grid features per sample (16 levels x 2)
constexpr uint32_t kColIn = 32;    // 16 SH + 15 geo_feat + 1 zero pad (network_ff.py:43)

// shared memory: sigma-net weights W_0..W_{ns} then colour-net weights W_0..W_{nc} (8 KB per 64-row matrix, 2 KB for the
// 16-row output matrices, each 1024-byte aligned), X double buffer, activation ping-pong pair.
template <bool TRAIN>
__global__ void __launch_bounds__(128)
k_nerf_fwd(const __half* __restrict__ enc, const float* __restrict__ dirs, const __half* __restrict__ w_sigma,
           const __half* __restrict__ w_color, const uint32_t M, const uint32_t ns, const uint32_t nc, const float density_scale,
           __half* __restrict__ fwd_buf, __half* __restrict__ color_in, __half* __restrict__ h0_out, float* __restrict__ sigmas,
           float* __restrict__ rgbs, const uint32_t ntiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    uint8_t* sWs = sm;                                  // ns matrices of 8 KB + 2 KB
    uint8_t* sWc = sWs + ns * kWBytes + 2048;           // nc matrices of 8 KB + 2 KB
    uint8_t* sX = sWc + nc * kWBytes + 2048;
    uint8_t* sA = sX + 2 * kTileBytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sA + 2 * kTileBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t row = tid;

    // weights of both nets + the first tile, all in flight together; TMEM allocation overlaps the copies
    load_rows_async(smem_u32(sWs), w_sigma, 64, kEncDim, tid);
    for (uint32_t m = 1; m < ns; m++) load_rows_async(smem_u32(sWs + m * kWBytes), w_sigma + 64 * kEncDim + (m - 1) * 4096, 64, 64, tid);
    load_rows_async(smem_u32(sWs + ns * kWBytes), w_sigma + 64 * kEncDim + (ns - 1) * 4096, 16, 64, tid);
    load_rows_async(smem_u32(sWc), w_color, 64, kColIn, tid);
    for (uint32_t m = 1; m < nc; m++) load_rows_async(smem_u32(sWc + m * kWBytes), w_color + 64 * kColIn + (m - 1) * 4096, 64, 64, tid);
    load_rows_async(smem_u32(sWc + nc * kWBytes), w_color + 64 * kColIn + (nc - 1) * 4096, 16, 64, tid);
    load_rows_async(smem_u32(sX), enc + (size_t)blockIdx.x * kRows * kEncDim, kRows, kEncDim, tid);
    cp_async_commit();
    if (warp == 0) tmem_alloc(tslot, 64);
    if (tid == 32) { mbar_init(mbar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0, it = 0;
    const uint32_t nsteps = ns + 1 + nc + 1;  // matmuls per sample

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const size_t r0 = (size_t)tile * kRows;
        uint8_t* X = sX + (it & 1u) * kTileBytes;
        cp_async_wait_all();
        fence_proxy_async();
        __syncthreads();  // X (and on the first pass the weights) landed; the previous tile is completely done
        if (tile + gridDim.x < ntiles) {
            load_rows_async(smem_u32(sX + ((it + 1u) & 1u) * kTileBytes), enc + (size_t)(tile + gridDim.x) * kRows * kEncDim, kRows, kEncDim, tid);
            cp_async_commit();
        }
        // this sample's direction: needed by the sigma epilogue, ns + 1 matmuls from now
        const float dx = __ldcs(dirs + (r0 + row) * 3), dy = __ldcs(dirs + (r0 + row) * 3 + 1), dz = __ldcs(dirs + (r0 + row) * 3 + 2);

        // step s: 0..ns = sigma net (input X), ns+1..ns+1+nc = colour net (input = the assembled colour row)
        // operand tiles: step 0 reads X; step s > 0 reads sA[(s-1)&1]; its epilogue writes sA[s&1]
        for (uint32_t s = 0; s < nsteps; s++) {
            const bool sig = s <= ns;
            const uint32_t m = sig ? s : s - (ns + 1);            // layer index inside its net
            const uint32_t nl = sig ? ns : nc;
            const bool last = m == nl;                            // the 16-wide output layer
            uint8_t* cur = s == 0 ? X : sA + ((s - 1u) & 1u) * kTileBytes;
            uint8_t* nxt = sA + (s & 1u) * kTileBytes;
            if (tid == 0) {
                tc_fence_after();
                const uint32_t K = m == 0 ? 32u : 64u, N = last ? 16u : 64u;
                const uint32_t idesc = make_idesc(128, N, false, false);
                const uint64_t a = desc_sw128(smem_u32(cur), 16), b = desc_sw128(smem_u32((sig ? sWs : sWc) + m * kWBytes), 16);
                for (uint32_t k = 0; k < K / 16; k++) umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, k > 0);
                umma_commit(mbar);
            }
            if (TRAIN && s > 0) {  // save the operand of this matmul while the tensor core works
                if (s == ns + 1) {  // the colour net's input rows (64 B each)
                    uint4* dst = reinterpret_cast<uint4*>(color_in + r0 * kColIn);
#pragma unroll
                    for (uint32_t i = 0; i < 4; i++) {
                        const uint32_t c = tid + i * 128;
                        __stcs(dst + c, *reinterpret_cast<const uint4*>(cur + sw128(c >> 2, c & 3u)));
                    }
                } else {  // hidden activations: sigma H_0..H_{ns-1} then colour H_0..H_{nc-1}
                    const uint32_t slot = sig ? s - 1 : s - 2;
                    uint4* dst = reinterpret_cast<uint4*>(fwd_buf + ((size_t)slot * M + r0) * 64);
#pragma unroll
                    for (uint32_t i = 0; i < 8; i++) {
                        const uint32_t c = tid + i * 128;
                        __stcs(dst + c, *reinterpret_cast<const uint4*>(cur + sw128(c >> 3, c & 7u)));
                    }
                }
            }
            mbar_wait(mbar, phase);
            phase ^= 1u;
            tc_fence_after();
            if (!last) {
                uint32_t r[64];
                tmem_ld32_nowait(taddr, r);
                tmem_ld32_nowait(taddr + 32, r + 32);
                tmem_wait_ld();
#pragma unroll
                for (uint32_t q = 0; q < 8; q++) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = fmaxf(__uint_as_float(r[q * 8 + i]), 0.0f);
                    *reinterpret_cast<uint4*>(nxt + sw128(row, q)) = pack8(v);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncthreads();
            } else if (sig) {
                // sigma epilogue: h (fp16) -> sigma = density_scale * exp(h0); colour row = [SH(dir) | h[1..15] | 0]
                float h[16], sh[16], c2[16];
                tmem_ld16(taddr, h);
                const __half h0 = __float2half_rn(h[0]);
                __stcs(sigmas + r0 + row, density_scale * expf(__half2float(h0)));
                if (TRAIN) h0_out[r0 + row] = h0;
                sh_basis(dx, dy, dz, 4, sh);
#pragma unroll
                for (int i = 0; i < 15; i++) c2[i] = h[i + 1];
                c2[15] = 0.0f;
                *reinterpret_cast<uint4*>(nxt + sw128(row, 0)) = pack8(sh);
                *reinterpret_cast<uint4*>(nxt + sw128(row, 1)) = pack8(sh + 8);
                *reinterpret_cast<uint4*>(nxt + sw128(row, 2)) = pack8(c2);
                *reinterpret_cast<uint4*>(nxt + sw128(row, 3)) = pack8(c2 + 8);
                tc_fence_before();
                fence_proxy_async();
                __syncthreads();
            } else {
                // colour epilogue: rgb = sigmoid(h[0..2]) evaluated on the fp16 output, result rounded to fp16 (torch.sigmoid on half)
                float h[16];
                tmem_ld16(taddr, h);
                float* o = rgbs + (r0 + row) * 3;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float x = __half2float(__float2half_rn(h[c]));
                    __stcs(o + c, __half2float(__float2half_rn(1.0f / (1.0f + expf(-x)))));
                }
                tc_fence_before();
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static size_t nerf_fwd_smem(uint32_t ns, uint32_t nc) { return 1024 + (ns + nc) * kWBytes + 4096 + 4 * kTileBytes + 64; }

}  // namespace lnrf

using namespace lnrf;

extern "C" {

size_t lnrf_nerf_wgrad_scratch_bytes(uint32_t num_layers_sigma, uint32_t num_layers_color) {
    return lnrf_ffmlp_wgrad_scratch_bytes(kEncDim, 16, 64, num_layers_sigma) + lnrf_ffmlp_wgrad_scratch_bytes(kColIn, 16, 64, num_layers_color);
}

int lnrf_nerf_forward(const void* enc_f16, const float* dirs, const void* w_sigma_f16, const void* w_color_f16, uint32_t M,
                      uint32_t num_layers_sigma, uint32_t num_layers_color, float density_scale, int train,
                      void* forward_buffer_f16, void* color_in_f16, void* h0_f16, float* sigmas, float* rgbs, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    LNRF_REQUIRE(M % 128 == 0, "nerf_forward: the sample count must be 128 * m, but got %u", M);
    LNRF_REQUIRE(ns >= 2 && nc >= 2 && ns <= kMaxLayers && nc <= kMaxLayers, "nerf_forward: num_layers outside [2, %u]", kMaxLayers);
    if (M == 0) return LNRF_OK;
    LNRF_REQUIRE(enc_f16 && dirs && w_sigma_f16 && w_color_f16 && sigmas && rgbs, "nerf_forward: null pointer");
    LNRF_REQUIRE(!train || (forward_buffer_f16 && color_in_f16 && h0_f16), "nerf_forward: training needs the three save buffers");
    LNRF_REQUIRE(((reinterpret_cast<uintptr_t>(enc_f16) | reinterpret_cast<uintptr_t>(w_sigma_f16) | reinterpret_cast<uintptr_t>(w_color_f16) |
                   reinterpret_cast<uintptr_t>(forward_buffer_f16) | reinterpret_cast<uintptr_t>(color_in_f16)) & 15) == 0,
                 "nerf_forward: tensors must be 16-byte aligned");
    const size_t smem = nerf_fwd_smem(ns, nc);
    LNRF_REQUIRE(smem <= 227 * 1024, "nerf_forward: networks need %zu B of shared memory (> 227 KiB)", smem);
    auto kern = train ? k_nerf_fwd<true> : k_nerf_fwd<false>;
    static std::atomic<size_t> s_max_smem[2] = {{0}, {0}};
    std::atomic<size_t>& mx = s_max_smem[train ? 1 : 0];
    if (smem > mx.load(std::memory_order_relaxed)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "nerf_forward");
        mx.store(smem, std::memory_order_relaxed);
    }
    const uint32_t ntiles = M / kRows;
    uint32_t per_sm = (uint32_t)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    const uint32_t cap = (uint32_t)kNumSMs * per_sm;
    const uint32_t grid = ntiles < cap ? ntiles : cap;
    kern<<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __half*)enc_f16, dirs, (const __half*)w_sigma_f16, (const __half*)w_color_f16, M, ns, nc, density_scale,
        (__half*)forward_buffer_f16, (__half*)color_in_f16, (__half*)h0_f16, sigmas, rgbs, ntiles);
    LNRF_LAUNCH_CHECK("nerf_forward");
    return LNRF_OK;
}

int lnrf_nerf_backward(const float* grad_sigmas, const float* grad_rgbs, const float* rgbs, const void* h0_f16, const void* enc_f16,
                       const void* color_in_f16, const void* w_sigma_f16, const void* w_color_f16, const void* forward_buffer_f16,
                       uint32_t M, uint32_t num_layers_sigma, uint32_t num_layers_color, float density_scale, void* grad_enc_f16,
                       void* grad_w_sigma_f16, void* grad_w_color_f16, int accumulate_wgrad, void* dh_scratch_f16,
                       void* wgrad_scratch, size_t wgrad_scratch_bytes, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    MlpShape shs, shc;
    if (int e = mlp_shape("nerf_backward", M, kEncDim, 16, ns, &shs)) return e;
    if (int e = mlp_shape("nerf_backward", M, kColIn, 16, nc, &shc)) return e;
    LNRF_REQUIRE(grad_sigmas && grad_rgbs && rgbs && h0_f16 && dh_scratch_f16 && grad_enc_f16, "nerf_backward: null pointer");
    const size_t need_s = lnrf_ffmlp_wgrad_scratch_bytes(kEncDim, 16, 64, ns), need_c = lnrf_ffmlp_wgrad_scratch_bytes(kColIn, 16, 64, nc);
    if (!wgrad_scratch || wgrad_scratch_bytes < need_s + need_c) {
        set_error("nerf_backward: wgrad scratch too small (%zu < %zu)", wgrad_scratch_bytes, need_s + need_c);
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    BwdGlue g{grad_rgbs, rgbs, grad_sigmas, (const __half*)h0_f16, density_scale, (__half*)dh_scratch_f16};
    const __half* fb = (const __half*)forward_buffer_f16;
    // colour net: dL/drgb -> dL/dh (glue fused), dW_color
    if (int e = ffmlp_bwd_run("nerf_backward(color)", nullptr, color_in_f16, w_color_f16, fb + (size_t)ns * M * 64, M, shc, 1, nullptr,
                              grad_w_color_f16, (uint8_t*)wgrad_scratch + need_s, need_c, &g, accumulate_wgrad, st))
        return e;
    // sigma net: dL/dh -> dL/denc, dW_sigma
    return ffmlp_bwd_run("nerf_backward(sigma)", dh_scratch_f16, enc_f16, w_sigma_f16, fb, M, shs, 1, grad_enc_f16, grad_w_sigma_f16,
                         wgrad_scratch, need_s, nullptr, accumulate_wgrad, st);
}

}  // extern "C"
