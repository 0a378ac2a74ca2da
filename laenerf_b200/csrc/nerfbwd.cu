// nerfbwd.cu -- backward of NeRFNetwork.forward (nerf/network_ff.py:51-79) as ONE warp-specialised sm_100a kernel that
// RECOMPUTES the hidden activations instead of reading them back (round 2; replaces the two k_mlp_bwd2 launches of round 1 and,
// with them, the reference's kernel_mlp_fused_backward + CUTLASS split-K GEMMs, ffmlp/src/ffmlp.cu:410-518, 783-887).
//
// Why.  The round-1 forward saved forward_buffer [ns+nc, M, 64] + color_in [M, 32] (142 MB written per step at M = 228 k) and the
// two backward launches read them back (217 MB): the two largest DRAM streams of the step after Adam, for kernels whose tensor
// work is ~14 us.  The hidden activations are a pure function of enc [M,32] (64 B/sample) and of h = sigma_net(enc) [M,16]
// (32 B/sample, kept by the forward: it is also needed for trunc_exp's backward), so the backward rebuilds them on the tensor
// cores -- 5 extra 128x64x64 layer steps per tile, ~900 tensor-pipe cycles -- and touches ~170 B/sample of DRAM instead of ~1600.
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the tile's enc rows ([M,32] fp16 through a 64-column SWIZZLE_128B box:
//               the out-of-bounds half is zero-filled, so the tile lands in the same 128-byte-row layout as every other operand)
//   warp 1      MMA issuer: one elected lane issues every tcgen05.mma of BOTH tile sets in a fixed interleaved order and
//               tcgen05.commit's to the set's mbarrier
//   warp 2      TMEM allocation (512 columns); warp 3 idle (epilogue warps must start at a multiple of 4: TMEM lane quarters)
//   warps 4-7, 8-11   two epilogue warpgroups, one per tile set: thread r owns row r of its tile = TMEM lane r
// Per 128-row tile the chain is 2(nc+ns)+2 steps (12 for LAENeRF's nets), each "MMA batch -> commit -> epilogue":
//   colour net   forward  cin -> C0 -> ... -> C(nc-1)                     (cin = [SH(dir) | h[1..15] | 0] rebuilt from the saved h)
//                backward G(nc)=dY;  m = nc..1: G(m-1) = (G(m) W(m)) .* relu'(C(m-1)) written over C(m-1), dW(m) += C(m-1)^T G(m)
//                m = 0:   dcin = G(0) W(0) (only columns 16..30 = dL/dgeo_feat are used), dW(0) += G(0)^T cin
//   sigma net    forward  enc -> H0 -> ... -> H(ns-1)  (tiles of the colour net reused), backward from dh = [dsigma * density_scale *
//                exp(clamp(h0)), dL/dgeo_feat] the same way; denc = G(0) W(0) goes to global memory as fp16
// The SAME weight tile W(m) [out rows x in cols] serves the forward (K-major B operand, K = in) and the dgrad (MN-major B operand,
// N = in, K = out): no transposed copies.  Weight gradients of both tile sets accumulate into ONE set of TMEM accumulators (the
// tensor pipe executes MMAs in issue order), flushed once per CTA into a per-CTA slice; k_wgrad_reduce adds the slices in a
// fixed order (deterministic).  Shared memory: weights 44 KB + 2 sets x (nc + 2) tiles x 16 KB = 204 KB.
#include "mlp_core.cuh"
#include "sh_core.cuh"
#include <string.h>

namespace lnrf {

constexpr uint32_t kBG = 2;                          // tile sets = epilogue warpgroups per CTA
constexpr uint32_t kBwdThreads = 128 * (1 + kBG);    // warps 0-3 control, then kBG x 4 epilogue warps
constexpr uint32_t kEnc = 32, kCin = 32;

struct NerfBwdArgs {
    const float* dirs;          // [M,3]
    const __half* h;            // [M,16] sigma-net output saved by the forward (h0 = log-density, h[1..15] = geo_feat)
    const float* grad_sigma;    // [M]
    const float* grad_rgb;      // [M,3]
    const float* rgb;           // [M,3] saved sigmoid outputs (fp16 values held in fp32)
    const __half* w_sigma;
    const __half* w_color;
    __half* grad_enc;           // [M,32]
    float* wgrad_sigma;         // [grid][n_params_sigma] per-CTA partial sums
    float* wgrad_color;         // [grid][n_params_color]
    const int* M_dev;           // optional device-side sample count (rows beyond it are padding: skipped)
    long long* dbg;             // optional [nsteps][2] cycle sums (block 0, set 0): waiting for the MMA batch / epilogue work (diagnostics)
    uint32_t M, ns, nc, ntiles;
    float density_scale;
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking: has the phase with this parity completed?
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_tile(const CUtensorMap* map, uint32_t smem_tile, uint64_t* bar, int32_t col, int32_t row) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_tile),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row)
                 : "memory");
}

// TMEM columns: [0, 64 kBG) working accumulators of the sets; then the weight-gradient accumulators of the colour net and of the
// sigma net: dW(n) (16 columns), dW(n-1) .. dW(1) (64 each), dW(0) (32).
__device__ __host__ inline uint32_t wacc_size(uint32_t n) { return 16u + 64u * (n - 1u) + 32u; }
__device__ __forceinline__ uint32_t wacc_col(uint32_t n, uint32_t m) {
    if (m == n) return 0u;
    if (m == 0u) return 16u + 64u * (n - 1u);
    return 16u + 64u * (n - 1u - m);
}

// Measured and dropped (scripts/diag_overlap.py): capping the kernel at 128 registers (launch bound 512) so that a CTA of the
// hash-grid backward fits beside it on every SM.  Launched together on two streams the pair took 164 us against 100 + 97 us one after
// the other -- the encoder kernel only gets 8 warps per SM while this one is resident -- and the spills cost this kernel 10 us.
#ifndef LNRF_BWD_LB
#define LNRF_BWD_LB kBwdThreads
#endif
__global__ void __launch_bounds__(LNRF_BWD_LB, 1)
k_nerf_bwd(const __grid_constant__ CUtensorMap tm_enc, const NerfBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    const uint32_t ns = a.ns, nc = a.nc;
    uint32_t ntiles = a.ntiles;
    pdl_trigger();
    pdl_wait();  // nothing below may run before the kernels ahead on the stream are complete (common.cuh)
    if (a.M_dev) {
        const uint32_t live = div_up((uint32_t)max(*a.M_dev, 0), kRows);
        ntiles = live < ntiles ? live : ntiles;
    }
    uint8_t* sWs = sm;                                   // ns matrices of 8 KB + 2 KB (16-row output matrix)
    uint8_t* sWc = sWs + ns * kWBytes + 2048;            // nc matrices of 8 KB + 2 KB
    uint8_t* sSets = sWc + nc * kWBytes + 2048;          // kBG sets of (nc + 2) tiles: C(0..nc-1), CIN, DY
    const uint32_t set_bytes = (nc + 2u) * kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSets + kBG * set_bytes);
    uint64_t* full = bars;               // [kBG] MMA batch of a step complete (tcgen05.commit)
    uint64_t* ready = bars + kBG;        // [kBG] the epilogue wrote the next operand (128 arrivals)
    uint64_t* xfull = bars + 2 * kBG;    // [kBG] enc tile landed (TMA complete_tx)
    uint64_t* dyfree = bars + 3 * kBG;   // [kBG] the DY tile may be overwritten by the enc rows (tcgen05.commit)
    uint64_t* wdone = bars + 4 * kBG;    // [kBG] the weight-gradient MMAs of a backward step are complete: its operand tiles may be overwritten
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 5 * kBG);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t stride = gridDim.x * kBG;
    const uint32_t nsteps = 2u * (nc + ns) + 2u;
    const uint32_t sB = 2u * nc + 1u;                    // first step of the sigma phase

    // ---- prologue: weights (all threads, cp.async), barriers, TMEM ----
    load_rows_async_n(smem_u32(sWs), a.w_sigma, 64, kEnc, tid, kBwdThreads);
    for (uint32_t m = 1; m < ns; m++) load_rows_async_n(smem_u32(sWs + m * kWBytes), a.w_sigma + 64 * kEnc + (m - 1) * 4096, 64, 64, tid, kBwdThreads);
    load_rows_async_n(smem_u32(sWs + ns * kWBytes), a.w_sigma + 64 * kEnc + (ns - 1) * 4096, 16, 64, tid, kBwdThreads);
    load_rows_async_n(smem_u32(sWc), a.w_color, 64, kCin, tid, kBwdThreads);
    for (uint32_t m = 1; m < nc; m++) load_rows_async_n(smem_u32(sWc + m * kWBytes), a.w_color + 64 * kCin + (m - 1) * 4096, 64, 64, tid, kBwdThreads);
    load_rows_async_n(smem_u32(sWc + nc * kWBytes), a.w_color + 64 * kCin + (nc - 1) * 4096, 16, 64, tid, kBwdThreads);
    cp_async_commit();
    if (warp == 2) tmem_alloc(tslot, 512);
    if (tid == 0) {
        for (uint32_t g = 0; g < kBG; g++) {
            mbar_init(full + g, 1);
            mbar_init(ready + g, 128);
            mbar_init(xfull + g, 1);
            mbar_init(dyfree + g, 1);
            mbar_init(wdone + g, 1);
        }
        fence_mbar_init();
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t wcol_c = tbase + 64u * kBG, wcol_s = wcol_c + wacc_size(nc);

    if (warp == 0) {
        // ===== TMA producer: the enc rows of every tile go into the set's DY tile once the colour backward has consumed dY =====
        if (lane_id() == 0) {
            uint32_t ph = 0;
            for (uint32_t it = 0;; it++) {
                bool any = false;
                for (uint32_t g = 0; g < kBG; g++) {
                    const uint32_t tile = blockIdx.x * kBG + g + it * stride;
                    if (tile >= ntiles) continue;
                    any = true;
                    mbar_wait_hot(dyfree + g, ph);
                    mbar_expect_tx(xfull + g, kTileBytes);
                    tma_load_tile(&tm_enc, smem_u32(sSets + g * set_bytes + (nc + 1u) * kTileBytes), xfull + g, 0, (int32_t)(tile * kRows));
                }
                if (!any) break;
                ph ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // Fixed interleaved order (set 0, set 1, set 0, ...): the mbarrier waits suspend in hardware and wake ~60 cycles after the
        // arrival, where a polling loop that serves "whichever set is ready" measured 20 % slower (poll + bookkeeping in local memory).
        // A backward step issues the dgrad first and commits it alone (`full`): the epilogue starts reading TMEM while the 8
        // weight-gradient MMAs of the step still run; their completion is a second commit (`wdone`) that the epilogue waits for
        // before it overwrites the tiles they read.
        uint32_t ph_ready = 0u, ph_x = 0u;   // both sets run the same step sequence: one parity each serves both
        for (uint32_t it = 0;; it++) {
            const bool v0 = blockIdx.x * kBG + it * stride < ntiles, v1 = blockIdx.x * kBG + 1u + it * stride < ntiles;
            if (!v0 && !v1) break;
            for (uint32_t s = 0; s < nsteps; s++) {
                // which net, which of its phases
                const bool sig = s >= sB;
                const uint32_t n = sig ? ns : nc, ls = sig ? s - sB : s;        // layers, step inside the net's part
                const uint32_t sW = smem_u32(sig ? sWs : sWc), wcol = sig ? wcol_s : wcol_c;
#pragma unroll
                for (uint32_t g = 0; g < kBG; g++) {
                    if (!(g == 0u ? v0 : v1)) continue;
                    mbar_wait_hot(ready + g, ph_ready);
                    if (s == sB) mbar_wait_hot(xfull + g, ph_x);
                    tc_fence_after();
                    const uint32_t set = smem_u32(sSets + g * set_bytes);
                    const uint32_t tCIN = set + nc * kTileBytes, tDY = tCIN + kTileBytes;
                    const uint32_t work = tbase + 64u * g;
                    const bool accf = it > 0u || g > 0u;   // set 0 always owns the CTA's first tile
                    // tile of hidden layer k of this net: colour C(k) = tile k; sigma H(k) = tile nc-1-k; inputs / output-gradient tiles
                    auto hid = [&](uint32_t k) { return set + (sig ? (nc - 1u - k) : k) * kTileBytes; };
                    const uint32_t tIN = sig ? tDY : tCIN;      // X (enc rows) | cin
                    const uint32_t tGout = sig ? tCIN : tDY;    // dh | dY
                    if (elect_one()) {
                        if (ls < n) {                 // forward hidden layer k = ls
                            const uint32_t k = ls;
                            const uint64_t da = desc_sw128(k ? hid(k - 1u) : tIN, 16), db = desc_sw128(sW + k * kWBytes, 16);
                            const uint32_t idesc = make_idesc(128, 64, false, false);
                            if (k == 0u) umma_chain<2>(work, da, db, 2, 2, idesc, false);
                            else umma_chain<4>(work, da, db, 2, 2, idesc, false);
                            umma_commit(full + g);
                        } else if (ls < 2u * n) {     // backward through matmul m = n .. 1
                            const uint32_t m = 2u * n - ls;
                            const uint32_t tG = m == n ? tGout : hid(m);
                            const uint64_t wa = desc_sw128(hid(m - 1u), kTileBytes), wb = desc_sw128(tG, kTileBytes);
                            const uint64_t da = desc_sw128(tG, 16), db = desc_sw128(sW + m * kWBytes, kTileBytes);
                            const uint32_t didesc = make_idesc(128, 64, false, true);
                            if (m == n) umma_chain<1>(work, da, db, 2, 128, didesc, false);     // K = 16 output channels
                            else umma_chain<4>(work, da, db, 2, 128, didesc, false);           // K = 64
                            umma_commit(full + g);
                            umma_chain<8>(wcol + wacc_col(n, m), wa, wb, 128, 128, make_idesc(128, m == n ? 16u : 64u, true, true), accf);
                            umma_commit(wdone + g);
                            if (s == nc) umma_commit(dyfree + g);   // dY consumed (dgrad + wgrad of the colour output layer)
                        } else {                      // input layer: dX = G(0) W(0) (N = 32), dW(0) += G(0)^T X
                            const uint64_t wa = desc_sw128(hid(0), kTileBytes), wb = desc_sw128(tIN, kTileBytes);
                            const uint64_t da = desc_sw128(hid(0), 16), db = desc_sw128(sW, kTileBytes);
                            umma_chain<4>(work, da, db, 2, 128, make_idesc(128, 32, false, true), false);
                            umma_commit(full + g);
                            umma_chain<8>(wcol + wacc_col(n, 0), wa, wb, 128, 128, make_idesc(128, 32, true, true), accf);
                            umma_commit(wdone + g);
                        }
                    }
                    __syncwarp();
                }
                ph_ready ^= 1u;
                if (s == sB) ph_x ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warpgroups =====
        const uint32_t g = (warp >> 2) - 1u, row = tid & 127u;
        uint8_t* set = sSets + g * set_bytes;
        uint8_t* tCIN = set + nc * kTileBytes;
        uint8_t* tDY = tCIN + kTileBytes;
        const uint32_t taddr = tbase + 64u * g + ((uint32_t)((warp & 3u) * 32u) << 16);
        uint32_t ph_full = 0, ph_w = 0;
        const uint32_t first = blockIdx.x * kBG + g;

        // per-row inputs of a tile: h (16 halves), direction, dL/drgb, rgb, dL/dsigma
        uint4 hq[2];
        float dir[3], grgb[3], srgb[3], gsig = 0.f;
        auto fetch = [&](uint32_t tile) {
            const size_t r = (size_t)tile * kRows + row;
            const uint4* hp = reinterpret_cast<const uint4*>(a.h + r * 16);
            hq[0] = __ldcs(hp); hq[1] = __ldcs(hp + 1);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                dir[c] = __ldcs(a.dirs + r * 3 + c);
                grgb[c] = __ldcs(a.grad_rgb + r * 3 + c);
                srgb[c] = __ldcs(a.rgb + r * 3 + c);
            }
            gsig = __ldcs(a.grad_sigma + r);
        };
        // cin row = [SH(dir) | h[1..15] | 0] and dY row = fp16 sigmoid backward of dL/drgb (3 real columns of 16)
        auto build_inputs = [&]() {
            float sh[16], hv[16];
            unpack8(hq[0], hv);
            unpack8(hq[1], hv + 8);
            sh_basis(dir[0], dir[1], dir[2], 4, sh);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 8; i++) pk[i] = pack_h2(sh[2 * i], sh[2 * i + 1]);
#pragma unroll
            for (int i = 0; i < 7; i++) pk[8 + i] = pack_h2(hv[2 * i + 1], hv[2 * i + 2]);
            pk[15] = pack_h2(hv[15], 0.0f);
#pragma unroll
            for (uint32_t q = 0; q < 4; q++)
                *reinterpret_cast<uint4*>(tCIN + sw128(row, q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            float v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float gg = __half2float(__float2half_rn(grgb[c]));  // the grad of the .float() cast of rgb
                v[c] = gg * (1.0f - srgb[c]) * srgb[c];
            }
            *reinterpret_cast<uint4*>(tDY + sw128(row, 0)) = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], 0.0f), 0u, 0u);
            *reinterpret_cast<uint4*>(tDY + sw128(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
        };
        auto publish = [&]() {   // operand tile written: visible to the async proxy, TMEM reads ordered, then tell the MMA warp
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(ready + g);
        };

        if (first < ntiles) {
            fetch(first);
            build_inputs();
            publish();
        }
        for (uint32_t tile = first; tile < ntiles; tile += stride) {
            const size_t r0 = (size_t)tile * kRows;
            const float my_gsig = gsig;
            const float my_h0 = __half2float(__ushort_as_half((unsigned short)(hq[0].x & 0xffffu)));
            const uint32_t next = tile + stride;
            if (next < ntiles) fetch(next);   // in flight during the whole chain
            for (uint32_t s = 0; s < nsteps; s++) {
                const bool sig = s >= sB;
                const uint32_t n = sig ? ns : nc, ls = sig ? s - sB : s;
                auto hid = [&](uint32_t k) { return set + (sig ? (nc - 1u - k) : k) * kTileBytes; };
                uint4 hrow[8];
                if (ls >= n && ls < 2u * n) {   // the saved activations this step masks with: stable, fetched while the tensor core works
                    const uint8_t* Hp = hid(2u * n - ls - 1u);
#pragma unroll
                    for (uint32_t q = 0; q < 8; q++) hrow[q] = *reinterpret_cast<const uint4*>(Hp + sw128(row, q));
                }
                const long long c0 = a.dbg ? clock64() : 0;
                mbar_wait_hot(full + g, ph_full);
                ph_full ^= 1u;
                tc_fence_after();
                const long long c1 = a.dbg ? clock64() : 0;
                if (ls < n) {
                    // forward hidden layer: ReLU, fp16, into the layer's tile
                    uint8_t* T = hid(ls);
                    uint32_t r[64];
                    tmem_ld32_nowait(taddr, r);
                    tmem_ld32_nowait(taddr + 32, r + 32);
                    tmem_wait_ld();
#pragma unroll
                    for (uint32_t q = 0; q < 8; q++) {
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) o[j] = pack_relu(__uint_as_float(r[q * 8 + 2 * j]), __uint_as_float(r[q * 8 + 2 * j + 1]));
                        *reinterpret_cast<uint4*>(T + sw128(row, q)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    publish();
                } else if (ls < 2u * n) {
                    // G(m-1) = dgrad .* relu'(H(m-1)), rounded to fp16, written over H(m-1)
                    const uint32_t m = 2u * n - ls;
                    uint8_t* Hp = hid(m - 1u);
                    uint32_t r[64];
                    tmem_ld32_nowait(taddr, r);
                    tmem_ld32_nowait(taddr + 32, r + 32);
                    tmem_wait_ld();
                    const __half2 zero2 = __float2half2_rn(0.0f);
                    uint32_t o[32];
#pragma unroll
                    for (uint32_t q = 0; q < 8; q++) {
                        const uint32_t hw[4] = {hrow[q].x, hrow[q].y, hrow[q].z, hrow[q].w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t pk = pack_h2(__uint_as_float(r[q * 8 + 2 * j]), __uint_as_float(r[q * 8 + 2 * j + 1]));
                            o[q * 4 + j] = pk & __hgt2_mask(*reinterpret_cast<const __half2*>(&hw[j]), zero2);
                        }
                    }
                    mbar_wait_hot(wdone + g, ph_w);   // the step's weight-gradient MMAs have read H(m-1): it may become G(m-1)
                    ph_w ^= 1u;
#pragma unroll
                    for (uint32_t q = 0; q < 8; q++)
                        *reinterpret_cast<uint4*>(Hp + sw128(row, q)) = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                    publish();
                } else if (!sig) {
                    // colour input layer done: dh = [dL/dsigma * density_scale * exp(clamp(h0, -15, 15)) | dL/dgeo_feat] into the CIN tile
                    float v[16], o[16];
                    tmem_ld16(taddr + 16, v);   // dL/dcin[:, 16:32] = dL/dgeo_feat (15) and the zero-pad column
                    o[0] = my_gsig * a.density_scale * expf(fminf(fmaxf(my_h0, -15.0f), 15.0f));
#pragma unroll
                    for (int i = 1; i < 16; i++) o[i] = v[i - 1];
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) pk[i] = pack_h2(o[2 * i], o[2 * i + 1]);
                    mbar_wait_hot(wdone + g, ph_w);   // dW(0) += G(0)^T cin has read the CIN tile
                    ph_w ^= 1u;
                    *reinterpret_cast<uint4*>(tCIN + sw128(row, 0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(tCIN + sw128(row, 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    publish();
                } else {
                    // sigma input layer done: dL/denc to global memory; then the next tile's inputs
                    __half* gi = a.grad_enc + (r0 + row) * kEnc;
#pragma unroll
                    for (uint32_t q = 0; q < 2; q++) {
                        float v[16];
                        tmem_ld16(taddr + q * 16, v);
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) pk[i] = pack_h2(v[2 * i], v[2 * i + 1]);
                        st_global_32B(gi + q * 16, pk);
                    }
                    mbar_wait_hot(wdone + g, ph_w);   // every MMA of this tile is complete: all of its tiles are free
                    ph_w ^= 1u;
                    if (next < ntiles) {
                        build_inputs();
                        publish();
                    } else {
                        tc_fence_before();
                    }
                }
                if (a.dbg && blockIdx.x == 0 && g == 0 && row == 0) {
                    a.dbg[2 * s] += c1 - c0;
                    a.dbg[2 * s + 1] += clock64() - c1;
                }
            }
        }
    }

    // ---- flush the weight-gradient accumulators: lanes 0..63 are real (64..127 hold the ignored second atom) ----
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 4 && warp < 6 && blockIdx.x * kBG < ntiles) {   // lane quarters 0 and 1 of the first epilogue warpgroup
        const uint32_t row = tid & 127u;
        const uint32_t lane_bits = (uint32_t)((warp & 3u) * 32u) << 16;
        for (int net = 0; net < 2; net++) {
            const uint32_t n = net ? nc : ns, in_dim = 32u;
            const uint32_t wcol = (net ? wcol_c : wcol_s) + lane_bits;
            const uint32_t nparams = 64u * (in_dim + 64u * (n - 1u) + 16u);
            float* slice = (net ? a.wgrad_color : a.wgrad_sigma) + (size_t)blockIdx.x * nparams;
            for (uint32_t m = 1; m <= n; m++) {   // acc_m[i][j] = dW_m[j][i]; lane = i -> coalesced over i
                const uint32_t n_m = m == n ? 16u : 64u;
                float* dst = slice + 64u * in_dim + (m - 1u) * 4096u;
                for (uint32_t q = 0; q < n_m / 16u; q++) {
                    float v[16];
                    tmem_ld16(wcol + wacc_col(n, m) + q * 16u, v);
#pragma unroll
                    for (int j = 0; j < 16; j++) dst[(q * 16u + j) * 64u + row] = v[j];
                }
            }
            for (uint32_t q = 0; q < in_dim / 16u; q++) {   // acc_0[j][i] = dW_0[j][i]; lane = j
                float v[16];
                tmem_ld16(wcol + wacc_col(n, 0) + q * 16u, v);
#pragma unroll
                for (int i = 0; i < 16; i++) slice[row * in_dim + q * 16u + i] = v[i];
            }
        }
    }
    if (blockIdx.x * kBG >= ntiles) {   // a CTA without a tile (device-side sample count below the capacity): its slices are zeros
        const uint32_t np_s = 64u * (32u + 64u * (ns - 1u) + 16u), np_c = 64u * (32u + 64u * (nc - 1u) + 16u);
        for (uint32_t i = tid; i < np_s; i += kBwdThreads) a.wgrad_sigma[(size_t)blockIdx.x * np_s + i] = 0.0f;
        for (uint32_t i = tid; i < np_c; i += kBwdThreads) a.wgrad_color[(size_t)blockIdx.x * np_c + i] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tbase, 512);
}

static size_t nerf_bwd_smem(uint32_t ns, uint32_t nc) { return 1024 + (ns + nc) * kWBytes + 4096 + kBG * (nc + 2) * kTileBytes + 256; }

}  // namespace lnrf

using namespace lnrf;

static long long* g_nerf_bwd_dbg = nullptr;  // diagnostics only (scripts/diag_bwd_steps.py): device buffer of per-step cycle sums

extern "C" {

LNRF_API void lnrf_debug_set_nerf_bwd_counters(long long* device_buffer) { g_nerf_bwd_dbg = device_buffer; }

int lnrf_nerf_backward_recompute_supported(uint32_t num_layers_sigma, uint32_t num_layers_color) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    return ns >= 2 && nc >= 2 && ns <= nc && nerf_bwd_smem(ns, nc) <= 227 * 1024 && 64u * kBG + wacc_size(nc) + wacc_size(ns) <= 512u;
}

int lnrf_nerf_backward_recompute(const float* grad_sigmas, const float* grad_rgbs, const float* rgbs, const void* h_f16, const void* enc_f16,
                                 const float* dirs, const void* w_sigma_f16, const void* w_color_f16, uint32_t M, const int32_t* M_dev,
                                 uint32_t num_layers_sigma, uint32_t num_layers_color, float density_scale, void* grad_enc_f16,
                                 void* grad_w_sigma_f16, void* grad_w_color_f16, int accumulate_wgrad, void* wgrad_scratch,
                                 size_t wgrad_scratch_bytes, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    LNRF_REQUIRE(M % 128 == 0, "nerf_backward_recompute: the sample count must be 128 * m, but got %u", M);
    if (!lnrf_nerf_backward_recompute_supported(ns, nc)) {
        set_error("nerf_backward_recompute: layer counts (%u, %u) outside what the recompute kernel is built for (2 <= sigma <= colour, "
                  "shared memory / TMEM budget); use lnrf_nerf_backward", ns, nc);
        return LNRF_ERR_UNSUPPORTED;
    }
    LNRF_REQUIRE(grad_sigmas && grad_rgbs && rgbs && h_f16 && enc_f16 && dirs && w_sigma_f16 && w_color_f16 && grad_enc_f16 && grad_w_sigma_f16 &&
                     grad_w_color_f16,
                 "nerf_backward_recompute: null pointer");
    LNRF_REQUIRE(((reinterpret_cast<uintptr_t>(enc_f16) | reinterpret_cast<uintptr_t>(h_f16) | reinterpret_cast<uintptr_t>(w_sigma_f16) |
                   reinterpret_cast<uintptr_t>(w_color_f16) | reinterpret_cast<uintptr_t>(grad_enc_f16)) & 31) == 0,
                 "nerf_backward_recompute: tensors must be 32-byte aligned");
    const size_t need_s = lnrf_ffmlp_wgrad_scratch_bytes(kEnc, 16, 64, ns), need_c = lnrf_ffmlp_wgrad_scratch_bytes(kCin, 16, 64, nc);
    if (!wgrad_scratch || wgrad_scratch_bytes < need_s + need_c) {
        set_error("nerf_backward_recompute: wgrad scratch too small (%zu < %zu)", wgrad_scratch_bytes, need_s + need_c);
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t np_s = 64u * (kEnc + 64u * (ns - 1u) + 16u), np_c = 64u * (kCin + 64u * (nc - 1u) + 16u);
    const uint32_t ntiles = M / kRows;
    const uint32_t want = div_up(ntiles, kBG);
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    if (M > 0) {
        const size_t smem = nerf_bwd_smem(ns, nc);
        static std::atomic<size_t> s_max{0};
        if (smem > s_max.load(std::memory_order_relaxed)) {
            cudaError_t e = cudaFuncSetAttribute(k_nerf_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(e, "nerf_backward_recompute");
            s_max.store(smem, std::memory_order_relaxed);
        }
        CUtensorMap tm;
        if (int e = make_tensor_map_2d(&tm, enc_f16, kEnc, M, kEnc * 2, 64, kRows, "nerf_backward_recompute")) return e;
        NerfBwdArgs a{};
        a.dirs = dirs; a.h = (const __half*)h_f16; a.grad_sigma = grad_sigmas; a.grad_rgb = grad_rgbs; a.rgb = rgbs;
        a.w_sigma = (const __half*)w_sigma_f16; a.w_color = (const __half*)w_color_f16; a.grad_enc = (__half*)grad_enc_f16;
        a.wgrad_sigma = (float*)wgrad_scratch; a.wgrad_color = (float*)((uint8_t*)wgrad_scratch + need_s);
        a.M_dev = M_dev; a.M = M; a.ns = ns; a.nc = nc; a.ntiles = ntiles; a.density_scale = density_scale;
        a.dbg = g_nerf_bwd_dbg;
        launch_pdl(k_nerf_bwd, grid, kBwdThreads, smem, st, tm, a);
        LNRF_LAUNCH_CHECK("nerf_backward_recompute");
    } else {
        cudaError_t e = cudaMemsetAsync(wgrad_scratch, 0, need_s + need_c, st);
        if (e != cudaSuccess) return cuda_fail(e, "nerf_backward_recompute");
    }
    // accumulate_wgrad bit 1: the caller runs the reduction itself (lnrf_nerf_wgrad_reduce), e.g. on another stream beside the encoder
    // backward, which does not depend on it
    if (accumulate_wgrad & 2) return LNRF_OK;
    return lnrf_nerf_wgrad_reduce(wgrad_scratch, wgrad_scratch_bytes, M, ns, nc, grad_w_sigma_f16, grad_w_color_f16, accumulate_wgrad & 1, stream);
}

int lnrf_nerf_wgrad_reduce(const void* wgrad_scratch, size_t wgrad_scratch_bytes, uint32_t M, uint32_t num_layers_sigma,
                           uint32_t num_layers_color, void* grad_w_sigma_f16, void* grad_w_color_f16, int accumulate, lnrf_stream_t stream) {
    const uint32_t ns = num_layers_sigma, nc = num_layers_color;
    LNRF_REQUIRE(wgrad_scratch && grad_w_sigma_f16 && grad_w_color_f16, "nerf_wgrad_reduce: null pointer");
    const size_t need_s = lnrf_ffmlp_wgrad_scratch_bytes(kEnc, 16, 64, ns), need_c = lnrf_ffmlp_wgrad_scratch_bytes(kCin, 16, 64, nc);
    LNRF_REQUIRE(wgrad_scratch_bytes >= need_s + need_c, "nerf_wgrad_reduce: wgrad scratch too small");
    const uint32_t np_s = 64u * (kEnc + 64u * (ns - 1u) + 16u), np_c = 64u * (kCin + 64u * (nc - 1u) + 16u);
    const uint32_t want = div_up(M / kRows, kBG);
    const uint32_t grid = want < (uint32_t)kNumSMs ? want : (uint32_t)kNumSMs;
    // with M_dev the number of CTAs that really had a tile is not known on the host: all `grid` slices are reduced (CTAs without a
    // tile write zeros into theirs)
    WgradPending ps{(const float*)wgrad_scratch, M > 0 ? grid : 1u, (__half*)grad_w_sigma_f16, np_s, accumulate & 1};
    WgradPending pc{(const float*)((const uint8_t*)wgrad_scratch + need_s), M > 0 ? grid : 1u, (__half*)grad_w_color_f16, np_c, accumulate & 1};
    return wgrad_reduce_pair("nerf_backward_recompute(wgrad)", ps, pc, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
