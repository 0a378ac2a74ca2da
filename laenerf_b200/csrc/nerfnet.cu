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
#include "render_core.cuh"
#include <string.h>

namespace lnrf {

constexpr uint32_t kEncDim = 32;   // hash-grid features per sample (16 levels x 2)
constexpr uint32_t kColIn = 32;    // 16 SH + 15 geo_feat + 1 zero pad (network_ff.py:43)

// ---- multi-tile-in-flight structure --------------------------------------------------------------------------------
// A layer step of one 128-row tile is a serial chain (MMA -> commit -> TMEM load -> ReLU/pack -> shared store -> fence ->
// barrier, ~1 us) that leaves the SM mostly idle, and shared memory (44 KB of weights per CTA) allowed only two CTAs per
// SM (ncu r1c: 12 % warps active, 9 % tensor pipe, long-scoreboard + barrier stalls).  So ONE CTA of 512 threads per SM
// runs kGroups = 4 independent tiles at once: warpgroup g (4 warps = the 128 TMEM lanes) owns tile slots, 64 TMEM columns,
// one mbarrier and one named barrier; the weights are shared.  Each group's chain runs IN PLACE in a single 16 KB tile
// (the MMA that read it has completed before the epilogue overwrites it) while the group's other tile receives the next
// input rows (cp.async prefetch), and the activations to be saved for the backward go to global memory straight from the
// epilogue registers as full 32-byte sectors (st.global.v8), not through a second pass over shared memory.
constexpr uint32_t kGroups = 4;

// shared memory: sigma-net weights W_0..W_{ns} then colour-net weights W_0..W_{nc} (8 KB per 64-row matrix, 2 KB for the
// 16-row output matrices, each 1024-byte aligned), then per group two 16 KB tiles, then kGroups mbarriers + the TMEM slot.
template <bool TRAIN>
__global__ void __launch_bounds__(128 * kGroups, 1)
k_nerf_fwd(const __half* __restrict__ enc, const float* __restrict__ dirs, const __half* __restrict__ w_sigma,
           const __half* __restrict__ w_color, const uint32_t M, const uint32_t ns, const uint32_t nc, const float density_scale,
           const __grid_constant__ CUtensorMap tm_fwd_buf, __half* __restrict__ color_in, __half* __restrict__ h0_out,
           float* __restrict__ sigmas, float* __restrict__ rgbs, uint32_t ntiles, const int* __restrict__ M_dev,
           const uint32_t sigma_only, __half* __restrict__ h_all_out) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    pdl_wait();  // nothing below may run before the kernels ahead on the stream are complete (common.cuh)
    if (M_dev) {  // device-side sample count: the render control block (row f-3) or the training marcher's counter; capacity = ntiles
        const uint32_t live = div_up((uint32_t)max(*M_dev, 0), kRows);
        ntiles = live < ntiles || ntiles == 0u ? live : ntiles;
        if (ntiles == 0u) return;
    }
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    uint8_t* sWs = sm;                                  // ns matrices of 8 KB + 2 KB
    uint8_t* sWc = sWs + ns * kWBytes + 2048;           // nc matrices of 8 KB + 2 KB
    uint8_t* sT = sWc + nc * kWBytes + 2048;            // kGroups x 2 tiles
    uint64_t* mbars = reinterpret_cast<uint64_t*>(sT + kGroups * 2 * kTileBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(mbars + kGroups);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, g = tid >> 7, gt = tid & 127u;
    const uint32_t row = gt;  // this thread's row of its group's tile == its TMEM lane (warp % 4 selects the 32-lane quarter)
    uint8_t* myT = sT + g * 2 * kTileBytes;
    uint64_t* mbar = mbars + g;
    const uint32_t stride = gridDim.x * kGroups;
    const uint32_t first = blockIdx.x * kGroups + g;

    // weights of both nets (all 512 threads) + every group's first tile, all in flight together; TMEM allocation overlaps
    load_rows_async_n(smem_u32(sWs), w_sigma, 64, kEncDim, tid, 128 * kGroups);
    for (uint32_t m = 1; m < ns; m++) load_rows_async_n(smem_u32(sWs + m * kWBytes), w_sigma + 64 * kEncDim + (m - 1) * 4096, 64, 64, tid, 128 * kGroups);
    load_rows_async_n(smem_u32(sWs + ns * kWBytes), w_sigma + 64 * kEncDim + (ns - 1) * 4096, 16, 64, tid, 128 * kGroups);
    if (!sigma_only) {
        load_rows_async_n(smem_u32(sWc), w_color, 64, kColIn, tid, 128 * kGroups);
        for (uint32_t m = 1; m < nc; m++) load_rows_async_n(smem_u32(sWc + m * kWBytes), w_color + 64 * kColIn + (m - 1) * 4096, 64, 64, tid, 128 * kGroups);
        load_rows_async_n(smem_u32(sWc + nc * kWBytes), w_color + 64 * kColIn + (nc - 1) * 4096, 16, 64, tid, 128 * kGroups);
    }
    if (first < ntiles) load_rows_async_n(smem_u32(myT), enc + (size_t)first * kRows * kEncDim, kRows, kEncDim, gt, 128);
    cp_async_commit();
    if (warp == 0) tmem_alloc(tslot, 64 * kGroups);
    if (tid == 32) {
        for (uint32_t i = 0; i < kGroups; i++) mbar_init(mbars + i, 1);
        fence_mbar_init();
    }
    float dx = 0.f, dy = 0.f, dz = 0.f;  // this sample's direction, fetched one tile ahead (the sigma epilogue needs it)
    if (first < ntiles && !sigma_only) {
        const float* d = dirs + ((size_t)first * kRows + row) * 3;
        dx = __ldcs(d); dy = __ldcs(d + 1); dz = __ldcs(d + 2);
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();   // weights + first tiles landed, TMEM allocated, mbarriers initialised
    tc_fence_after();
    const uint32_t tmem = *tslot + g * 64u;                                  // this group's 64 accumulator columns
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3u) * 32u) << 16);     // + this warp's lane quarter
    uint32_t phase = 0, it = 0;
    const uint32_t nsteps = sigma_only ? ns + 1 : ns + 1 + nc + 1;  // matmuls per sample (density(): the sigma net alone)

    for (uint32_t tile = first; tile < ntiles; tile += stride, it++) {
        const size_t r0 = (size_t)tile * kRows;
        uint8_t* T = myT + (it & 1u) * kTileBytes;   // this tile's chain buffer (its input rows are already here)
        if (it > 0) {
            cp_async_wait_all();
            fence_proxy_async();
            group_barrier(g);  // the prefetched rows landed; the previous tile of this group is completely done
        }
        const float cx = dx, cy = dy, cz = dz;
        if (tile + stride < ntiles) {
            load_rows_async_n(smem_u32(myT + ((it + 1u) & 1u) * kTileBytes), enc + (size_t)(tile + stride) * kRows * kEncDim, kRows, kEncDim, gt, 128);
            cp_async_commit();
            if (!sigma_only) {
                const float* d = dirs + ((size_t)(tile + stride) * kRows + row) * 3;
                dx = __ldcs(d); dy = __ldcs(d + 1); dz = __ldcs(d + 2);
            }
        }

        // step s: 0..ns = sigma net, ns+1..ns+1+nc = colour net; every step reads T and its epilogue rewrites T in place
        for (uint32_t s = 0; s < nsteps; s++) {
            const bool sig = s <= ns;
            const uint32_t m = sig ? s : s - (ns + 1);            // layer index inside its net
            const bool last = m == (sig ? ns : nc);               // the 16-wide output layer
            if ((gt >> 5) == 0) {  // the group's first warp, uniformly; one elected lane issues (see umma_chain)
                tc_fence_after();
                const uint32_t idesc = make_idesc(128, last ? 16u : 64u, false, false);
                const uint64_t a = desc_sw128(smem_u32(T), 16), b = desc_sw128(smem_u32((sig ? sWs : sWc) + m * kWBytes), 16);
                if (elect_one()) {
                    if (m == 0) umma_chain<2>(tmem, a, b, 2, 2, idesc, false);   // K = 32
                    else umma_chain<4>(tmem, a, b, 2, 2, idesc, false);          // K = 64
                    // the TMA store of the previous step's activations (issued below) reads T too: the commit that lets the
                    // epilogue overwrite T is issued only after that read has finished (it overlaps the MMAs)
                    if (TRAIN) tma_store_wait_read();
                    umma_commit(mbar);
                }
                __syncwarp();
            }
            mbar_wait_hot(mbar, phase);
            phase ^= 1u;
            tc_fence_after();
            if (!last) {
                uint32_t r[64], pk[32];
                tmem_ld32_nowait(taddr, r);
                tmem_ld32_nowait(taddr + 32, r + 32);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i++) pk[i] = pack_relu(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
#pragma unroll
                for (uint32_t q = 0; q < 8; q++)
                    *reinterpret_cast<uint4*>(T + sw128(row, q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                tc_fence_before();
                fence_proxy_async();
                group_barrier(g);
                if (TRAIN && (gt >> 5) == 0) {  // hidden activations: sigma H_0..H_{ns-1} then colour H_0..H_{nc-1}, one TMA store per tile
                    const uint32_t slot = sig ? s : s - 1;
                    if (elect_one()) {
                        tma_store_tile(&tm_fwd_buf, smem_u32(T), 0, (int32_t)(slot * M + (uint32_t)r0));
                        tma_store_commit();
                    }
                    __syncwarp();
                }
            } else if (sig) {
                // sigma epilogue: h (fp16) -> sigma = density_scale * exp(h0); colour row = [SH(dir) | h[1..15] | 0]
                float h[16], sh[16];
                uint32_t pk[16];
                tmem_ld16(taddr, h);
                const __half h0 = __float2half_rn(h[0]);
                if (sigma_only) {
                    __stcs(sigmas + r0 + row, density_scale * expf(__half2float(h0)));
                    tc_fence_before();
                    continue;
                }
                sh_basis(cx, cy, cz, 4, sh);
#pragma unroll
                for (int i = 0; i < 8; i++) pk[i] = pack_h2(sh[2 * i], sh[2 * i + 1]);
#pragma unroll
                for (int i = 0; i < 7; i++) pk[8 + i] = pack_h2(h[2 * i + 1], h[2 * i + 2]);
                pk[15] = pack_h2(h[15], 0.0f);
#pragma unroll
                for (uint32_t q = 0; q < 4; q++)
                    *reinterpret_cast<uint4*>(T + sw128(row, q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                tc_fence_before();
                fence_proxy_async();
                group_barrier(g);
                __stcs(sigmas + r0 + row, density_scale * expf(__half2float(h0)));
                if (h_all_out) {  // lean training forward: the 16 outputs of the sigma net are all the backward needs besides enc (nerfbwd.cu)
                    uint32_t ph[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) ph[i] = pack_h2(h[2 * i], h[2 * i + 1]);
                    st_global_32B(h_all_out + (r0 + row) * 16, ph);
                }
                if (TRAIN) {
                    h0_out[r0 + row] = h0;
                    __half* dst = color_in + (r0 + row) * kColIn;
                    st_global_32B(dst, pk);
                    st_global_32B(dst + 16, pk + 8);
                }
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
    if (TRAIN && (gt >> 5) == 0) {
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tslot, 64 * kGroups);
}

// ONE high-water mark per kernel instantiation, shared by every entry point that launches it: the attribute is per function, so
// separate marks let a later, smaller request (lnrf_nerf_density) lower it under an earlier, larger one (the render loop).
static int ensure_nerf_fwd_smem(bool train, size_t smem, const char* who) {
    static std::atomic<size_t> s_max_smem[2] = {{0}, {0}};
    std::atomic<size_t>& mx = s_max_smem[train ? 1 : 0];
    if (smem > mx.load(std::memory_order_relaxed)) {
        cudaError_t e = train ? cudaFuncSetAttribute(k_nerf_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(k_nerf_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, who);
        mx.store(smem, std::memory_order_relaxed);
    }
    return LNRF_OK;
}

static size_t nerf_fwd_smem(uint32_t ns, uint32_t nc) { return 1024 + (ns + nc) * kWBytes + 4096 + kGroups * 2 * kTileBytes + 128; }

int nerf_forward_dev_launch(const void* enc_f16, const float* dirs, const void* w_sigma_f16, const void* w_color_f16, uint32_t M_cap,
                            const int32_t* M_dev, uint32_t ns, uint32_t nc, float density_scale, float* sigmas, float* rgbs,
                            cudaStream_t st) {
    LNRF_REQUIRE(ns >= 2 && nc >= 2 && ns <= kMaxLayers && nc <= kMaxLayers, "render_rounds: num_layers outside [2, %u]", kMaxLayers);
    if (M_cap == 0) return LNRF_OK;
    LNRF_REQUIRE(enc_f16 && dirs && w_sigma_f16 && w_color_f16 && sigmas && rgbs && M_dev, "render_rounds(network): null pointer");
    const size_t smem = nerf_fwd_smem(ns, nc);
    LNRF_REQUIRE(smem <= 227 * 1024, "render_rounds: networks need %zu B of shared memory (> 227 KiB)", smem);
    if (int e = ensure_nerf_fwd_smem(false, smem, "render_rounds(network)")) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const uint32_t want = div_up(div_up(M_cap, kRows), kGroups);
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    launch_pdl(k_nerf_fwd<false>, grid, 128 * kGroups, smem, st, (const __half*)enc_f16, dirs, (const __half*)w_sigma_f16, (const __half*)w_color_f16, M_cap,
                                                         ns, nc, density_scale, tm, nullptr, nullptr, sigmas, rgbs, div_up(M_cap, kRows), M_dev, 0u, nullptr);
    LNRF_LAUNCH_CHECK("render_rounds(network)");
    return LNRF_OK;
}

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
    if (int e = ensure_nerf_fwd_smem(train != 0, smem, "nerf_forward")) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (train) {
        LNRF_REQUIRE((uint64_t)(ns + nc) * M < (1ull << 31), "nerf_forward: forward_buffer too large for one tensor map");
        if (int e = make_tile_tensor_map(&tm, forward_buffer_f16, (uint64_t)(ns + nc) * M, "nerf_forward")) return e;
    }
    const uint32_t ntiles = M / kRows;
    const uint32_t want = div_up(ntiles, kGroups);  // one persistent CTA per SM, kGroups tiles in flight each
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    launch_pdl(kern, grid, 128 * kGroups, smem, reinterpret_cast<cudaStream_t>(stream),
        (const __half*)enc_f16, dirs, (const __half*)w_sigma_f16, (const __half*)w_color_f16, M, ns, nc, density_scale,
        tm, (__half*)color_in_f16, (__half*)h0_f16, sigmas, rgbs, ntiles, nullptr, 0u, nullptr);
    LNRF_LAUNCH_CHECK("nerf_forward");
    return LNRF_OK;
}

int lnrf_nerf_forward_lean(const void* enc_f16, const float* dirs, const void* w_sigma_f16, const void* w_color_f16, uint32_t M,
                           const int32_t* M_dev, uint32_t num_layers_sigma, uint32_t num_layers_color, float density_scale, void* h_f16,
                           float* sigmas, float* rgbs, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    LNRF_REQUIRE(M % 128 == 0, "nerf_forward_lean: the sample count must be 128 * m, but got %u", M);
    LNRF_REQUIRE(ns >= 2 && nc >= 2 && ns <= kMaxLayers && nc <= kMaxLayers, "nerf_forward_lean: num_layers outside [2, %u]", kMaxLayers);
    if (M == 0) return LNRF_OK;
    LNRF_REQUIRE(enc_f16 && dirs && w_sigma_f16 && w_color_f16 && sigmas && rgbs, "nerf_forward_lean: null pointer");
    LNRF_REQUIRE(((reinterpret_cast<uintptr_t>(enc_f16) | reinterpret_cast<uintptr_t>(w_sigma_f16) | reinterpret_cast<uintptr_t>(w_color_f16)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(h_f16) & 31) == 0,
                 "nerf_forward_lean: tensors must be 16-byte (h: 32-byte) aligned");
    const size_t smem = nerf_fwd_smem(ns, nc);
    LNRF_REQUIRE(smem <= 227 * 1024, "nerf_forward_lean: networks need %zu B of shared memory (> 227 KiB)", smem);
    if (int e = ensure_nerf_fwd_smem(false, smem, "nerf_forward_lean")) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const uint32_t ntiles = M / kRows;
    const uint32_t want = div_up(ntiles, kGroups);
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    launch_pdl(k_nerf_fwd<false>, grid, 128 * kGroups, smem, reinterpret_cast<cudaStream_t>(stream),
        (const __half*)enc_f16, dirs, (const __half*)w_sigma_f16, (const __half*)w_color_f16, M, ns, nc, density_scale, tm, nullptr, nullptr,
        sigmas, rgbs, ntiles, M_dev, 0u, (__half*)h_f16);
    LNRF_LAUNCH_CHECK("nerf_forward_lean");
    return LNRF_OK;
}

int lnrf_nerf_density(const void* enc_f16, const void* w_sigma_f16, uint32_t M, uint32_t num_layers_sigma, float density_scale,
                      float* sigmas, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma;
    LNRF_REQUIRE(M % 128 == 0, "nerf_density: the sample count must be 128 * m, but got %u", M);
    LNRF_REQUIRE(ns >= 2 && ns <= kMaxLayers, "nerf_density: num_layers outside [2, %u]", kMaxLayers);
    if (M == 0) return LNRF_OK;
    LNRF_REQUIRE(enc_f16 && w_sigma_f16 && sigmas, "nerf_density: null pointer");
    const size_t smem = nerf_fwd_smem(ns, 2);  // the colour-net slots stay empty
    if (int e = ensure_nerf_fwd_smem(false, smem, "nerf_density")) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const uint32_t ntiles = M / kRows;
    const uint32_t want = div_up(ntiles, kGroups);
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    launch_pdl(k_nerf_fwd<false>, grid, 128 * kGroups, smem, reinterpret_cast<cudaStream_t>(stream),
        (const __half*)enc_f16, nullptr, (const __half*)w_sigma_f16, nullptr, M, ns, 0u, density_scale, tm, nullptr, nullptr, sigmas, nullptr,
        ntiles, nullptr, 1u, nullptr);
    LNRF_LAUNCH_CHECK("nerf_density");
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
    // colour net: dL/drgb -> dL/dh (glue fused), dW_color; sigma net: dL/dh -> dL/denc, dW_sigma.  Nothing reads the weight
    // gradients before the optimizer, so both fixed-order reductions of the per-CTA partial sums run as ONE launch at the end.
    WgradPending pc{}, ps{};
    if (int e = ffmlp_bwd_run("nerf_backward(color)", nullptr, color_in_f16, w_color_f16, fb + (size_t)ns * M * 64, M, shc, 1, nullptr,
                              grad_w_color_f16, (uint8_t*)wgrad_scratch + need_s, need_c, &g, accumulate_wgrad, st, &pc))
        return e;
    if (int e = ffmlp_bwd_run("nerf_backward(sigma)", dh_scratch_f16, enc_f16, w_sigma_f16, fb, M, shs, 1, grad_enc_f16, grad_w_sigma_f16,
                              wgrad_scratch, need_s, nullptr, accumulate_wgrad, st, &ps))
        return e;
    return wgrad_reduce_pair("nerf_backward(wgrad)", ps, pc, st);
}

}  // extern "C"
