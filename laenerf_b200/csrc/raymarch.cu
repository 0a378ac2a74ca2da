// raymarch.cu -- occupancy-grid ray marching and front-to-back compositing for sm_100a.
//
// Replaces raymarching/src/raymarching.cu of the reference (kernels :91-1142).  Design:
//   * marching: a GROUP of lanes per ray (32 for training, 8 for inference rounds) evaluates a window of
//     consecutive members of the ray's t-sequence in parallel -- occupancy-byte loads of a whole window are in
//     flight together instead of one dependent load per visited voxel -- and ballots resolve which members the
//     reference visits (march_core.cuh).  Training is ONE pass: sample parameters are parked in shared memory,
//     slots are handed out by a warp-level scan inside the block plus a decoupled look-back across blocks
//     (deterministic ray-id order; the reference uses two marches and global atomics), and the sample buffers are
//     written as flat, fully coalesced 16-byte vectors.
//   * compositing: warp per ray; transmittance is a multiplicative warp scan, colour/depth are warp reductions.
#include "common.cuh"
#include "march_core.cuh"
#include "render_core.cuh"

extern "C" {
static int check_march_common(uint32_t C, uint32_t H, uint32_t max_steps, const char* who);

}

// LNRF_MARCH_JUMP=0 keeps the serial window resolve (A/B measurement; results are identical either way)
static lnrf::MarchParams march_params_env(float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    const char* e = getenv("LNRF_MARCH_JUMP");  // read per call: tests flip it inside one process
    const int jump = e ? atoi(e) : 1;
    lnrf::MarchParams p = lnrf::make_march_params(bound, dt_gamma, max_steps, C, H);
    if (!jump) p.jump = 0;
    const char* f = getenv("LNRF_MARCH_FF");  // 0: no fast-forward over pending skips (A/B; identical results)
    if (f && atoi(f) == 0) p.fast_forward = 0;
    return p;
}

namespace lnrf {

// =========================================================================================================
// utilities (bit-exact integer / IEEE work)
// =========================================================================================================

// raymarching.cu:91-145
__global__ void __launch_bounds__(256) k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                  const float* __restrict__ aabb, uint32_t N, float min_near,
                                                  float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float rdx = f_div(1.0f, rays_d[n * 3]), rdy = f_div(1.0f, rays_d[n * 3 + 1]), rdz = f_div(1.0f, rays_d[n * 3 + 2]);
    const float a0 = __ldg(aabb), a1 = __ldg(aabb + 1), a2 = __ldg(aabb + 2), a3 = __ldg(aabb + 3), a4 = __ldg(aabb + 4), a5 = __ldg(aabb + 5);
    float near = f_mul(f_add(a0, -ox), rdx), far = f_mul(f_add(a3, -ox), rdx);
    if (near > far) { const float c = near; near = far; far = c; }
    float near_y = f_mul(f_add(a1, -oy), rdy), far_y = f_mul(f_add(a4, -oy), rdy);
    if (near_y > far_y) { const float c = near_y; near_y = far_y; far_y = c; }
    bool miss = (near > far_y || near_y > far);
    if (!miss) {
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = f_mul(f_add(a2, -oz), rdz), far_z = f_mul(f_add(a5, -oz), rdz);
        if (near_z > far_z) { const float c = near_z; near_z = far_z; far_z = c; }
        miss = (near > far_z || near_z > far);
        if (!miss) {
            if (near_z > near) near = near_z;
            if (far_z < far) far = far_z;
            if (near < min_near) near = min_near;
        }
    }
    if (miss) near = far = 3.402823466e+38f;  // std::numeric_limits<float>::max(), raymarching.cu:122
    nears[n] = near;
    fars[n] = far;
}

// raymarching.cu:162-198
__global__ void __launch_bounds__(256) k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      float radius, uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float RPI = 0.3183098861837907f;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[n * 2] = 2 * theta * RPI - 1;
    coords[n * 2 + 1] = phi * RPI;
}

// raymarching.cu:214-254
__global__ void __launch_bounds__(256) k_morton3d(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int)morton3d((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
__global__ void __launch_bounds__(256) k_morton3d_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[n * 3] = (int)morton3d_invert((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int)morton3d_invert((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int)morton3d_invert((uint32_t)(ind >> 2));
}

// raymarching.cu:267-289.  One thread packs 4 output bytes from 32 floats (8 x 16-byte loads), so a warp reads
// 4 KiB contiguous and writes 128 B contiguous.  N = number of output bytes.
__global__ void __launch_bounds__(256) k_packbits(const float* __restrict__ grid, uint32_t N, float thresh,
                                                  uint8_t* __restrict__ bitfield) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;  // output word index
    const uint32_t n0 = w * 4;
    if (n0 >= N) return;
    if (n0 + 4 <= N && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0) && ((reinterpret_cast<uintptr_t>(bitfield) & 3) == 0)) {
        const float4* g4 = reinterpret_cast<const float4*>(grid) + (size_t)w * 8;
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = __ldcs(g4 + i);
            const uint32_t nib = (v.x > thresh ? 1u : 0u) | (v.y > thresh ? 2u : 0u) | (v.z > thresh ? 4u : 0u) | (v.w > thresh ? 8u : 0u);
            word |= nib << (4 * i);
        }
        reinterpret_cast<uint32_t*>(bitfield)[w] = word;
    } else {
        for (uint32_t n = n0; n < N && n < n0 + 4; n++) {
            uint32_t bits = 0;
            for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (1u << i) : 0u;
            bitfield[n] = (uint8_t)bits;
        }
    }
}

// =========================================================================================================
// group marcher
// =========================================================================================================

template <int G>
struct Group {
    int gl, gshift;
    __device__ Group() {
        const int lane = threadIdx.x & 31;
        gl = lane & (G - 1);
        gshift = lane & ~(G - 1);
    }
    static constexpr unsigned kMask = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    __device__ __forceinline__ unsigned ballot(bool pred) const { return (__ballot_sync(kFull, pred) >> gshift) & kMask; }
    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(kFull, v, gshift + src); }
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(kFull, v, gshift + src); }
    __device__ __forceinline__ unsigned shfl(unsigned v, int src) const { return __shfl_sync(kFull, v, gshift + src); }
    static constexpr int kLog = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : (G == 4) ? 2 : 1;
};

__device__ __forceinline__ unsigned low_bits(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// Marches one ray per group until `max_emit` samples were visited or t >= far.  `emit(rank, s, dt, probe,
// prev_after)` is called by the lane that holds visited sample number `rank` (0-based); prev_after is the t after
// the previous emitted sample (or t_start), only maintained when TRACK_LAST.  With EDIT the probe carries the bit
// index so the caller can test a second bitfield.  Control flow is warp-uniform (groups that finished idle).
template <int G, bool TRACK_LAST, class Emit>
__device__ __forceinline__ uint32_t march_group(const Group<G>& grp, const MarchParams& p, const Ray& r,
                                                const uint8_t* __restrict__ grid, float t, float far, uint32_t max_emit,
                                                bool alive, Emit&& emit) {
    uint32_t cnt = 0;
    float pend = -INFINITY;  // target of a skip that ran past the previous window
    float last_after = t;
    while (__any_sync(kFull, alive)) {
        if (p.fast_forward) t = march_fast_forward(p, t, pend, G);  // straight to the member a pending skip lands on (group-uniform)
        float nxt;
        WindowInfo wi;
        const float s = march_window<G>(p, t, grp.gl, &nxt, &wi);
        const float dt = p.dt_const ? p.dt0 : march_dt(p, s);
        const bool valid = alive && (s < far);
        Probe q;
        q.x = q.y = q.z = 0.f; q.tt = 0.f; q.index = 0u;
        bool occ = false;
        if (valid) {
            q = march_probe(p, r, s, dt);
            occ = (__ldg(grid + (q.index >> 3)) >> (q.index & 7u)) & 1u;
        }
        const unsigned occm = grp.ballot(occ), valm = grp.ballot(valid);
        const unsigned reach = grp.ballot(s >= pend);
        unsigned vis = 0;
        int v = reach ? (__ffs(reach) - 1) : G;
        if (G == 32 && p.jump && __all_sync(kFull, !alive || wi.closed)) {  // narrow groups: the serial loop is shorter than log2(G) rounds + bookkeeping (measured: 20.7 vs 61.9 ms per 800x800 frame)
            // ---- jump-table resolve (every group of the warp holds a closed-form window; tests/test_march_core.py checks this
            // data flow lane by lane against the sequential marcher).  Successor of a lane: the next lane when its cell is
            // occupied (t += dt), the first member >= the voxel exit when it is empty, nothing when t >= far (the ray ends
            // there).  Pointer doubling then gives every lane its whole orbit in log2(G) shuffle rounds; the orbit of the first
            // visited lane is what the reference visits.  No data-dependent loop, no divergence.
            int nx = !valid ? G : (occ ? grp.gl + 1 : march_jump(wi, grp.gl, q.tt, G));
            unsigned orb = 1u << grp.gl;
#pragma unroll
            for (int rd = 0; rd < Group<G>::kLog; rd++) {
                const int src = nx < G ? nx : G - 1;
                const int n2 = grp.shfl(nx, src);
                const unsigned o2 = grp.shfl(orb, src);
                if (nx < G) { orb |= o2; nx = n2; }
            }
            const unsigned visited = grp.shfl(orb, v < G ? v : 0);
            const unsigned inval = visited & ~valm, visv = visited & valm;
            const unsigned emitm = visv & occm;
            const uint32_t room = max_emit - cnt;
            const int last = visv ? (31 - __clz(visv)) : 0;
            const float tt_last = grp.shfl(q.tt, last);
            // the first `room` samples when the budget ends the ray inside this window (num_steps < max_steps, raymarching.cu:359)
            const unsigned keep = grp.ballot(((emitm >> grp.gl) & 1u) && (uint32_t)__popc(emitm & low_bits(grp.gl)) < room);
            if (alive && v < G) {
                pend = -INFINITY;
                if (emitm != 0u && (uint32_t)__popc(emitm) >= room) {
                    vis = keep;
                    alive = false;
                } else {
                    vis = emitm;
                    if (inval) alive = false;                              // a visited member has t >= far
                    else if (!((occm >> last) & 1u)) pend = tt_last;       // the last skip runs past this window
                }
            }
        } else {
            if (v < G) pend = -INFINITY;
            bool res = alive && (v < G);
            while (__any_sync(kFull, res)) {
                const int vv = v < G ? v : (G - 1);
                const float tt = grp.shfl(q.tt, vv);
                const unsigned ge = grp.ballot(s >= tt);
                if (res) {
                    if (!((valm >> v) & 1u)) {  // t >= far at a visited member: the ray is finished
                        alive = false;
                        res = false;
                    } else if ((occm >> v) & 1u) {  // occupied: every following occupied lane is visited too (t += dt)
                        const unsigned rest = (~occm & Group<G>::kMask) >> v;
                        int run = rest ? (__ffs(rest) - 1) : (G - v);
                        const int room = (int)(max_emit - cnt) - __popc(vis);
                        if (run >= room) {  // reached the sample budget (num_steps < max_steps, raymarching.cu:359)
                            run = room;
                            alive = false;
                            res = false;
                        }
                        vis |= low_bits(run) << v;
                        v += run;
                        if (v >= G) res = false;
                    } else {  // empty: do { t += dt } while (t < tt)  ==> first later member with s >= tt
                        const unsigned m = ge & ~low_bits(v + 1) & Group<G>::kMask;
                        if (m) {
                            v = __ffs(m) - 1;
                        } else {
                            pend = tt;
                            v = G;
                            res = false;
                        }
                    }
                }
            }
        }
        // emit the visited samples of this window
        const float after = f_add(s, dt);
        float prev_after = last_after;
        if (TRACK_LAST) {
            const unsigned below = vis & low_bits(grp.gl);
            const int pl = below ? (31 - __clz(below)) : 0;
            const float pa = grp.shfl(after, pl);
            if (below) prev_after = pa;
            const int hl = vis ? (31 - __clz(vis)) : 0;
            const float la = grp.shfl(after, hl);
            if (vis) last_after = la;
        }
        if ((vis >> grp.gl) & 1u) emit(cnt + (uint32_t)__popc(vis & low_bits(grp.gl)), s, dt, q, prev_after);
        cnt += (uint32_t)__popc(vis);
        t = nxt;
    }
    return cnt;
}

// =========================================================================================================
// training march: one pass, deterministic slots
// =========================================================================================================

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Flat, coalesced write of `nflt` floats starting at float index `first` of `base` where element e is produced by
// f(e); uses 16-byte stores for the aligned body.  `base` must be 16-byte aligned (torch allocations are).
template <class F>
__device__ __forceinline__ void warp_write_flat(float* __restrict__ base, size_t first, uint32_t nflt, int lane, F&& f) {
    const uint32_t head0 = (uint32_t)((4 - (first & 3)) & 3);
    const uint32_t head = head0 < nflt ? head0 : nflt;
    if ((uint32_t)lane < head) st_cs(base + first + lane, f((uint32_t)lane));
    const uint32_t nvec = (nflt - head) >> 2;
    float4* vb = reinterpret_cast<float4*>(base + first + head);
    for (uint32_t v = lane; v < nvec; v += 32) {
        const uint32_t e = head + 4 * v;
        float4 o;
        o.x = f(e); o.y = f(e + 1); o.z = f(e + 2); o.w = f(e + 3);
        st_cs(vb + v, o);
    }
    const uint32_t done = head + 4 * nvec;
    if (done + lane < nflt) st_cs(base + first + done + lane, f(done + (uint32_t)lane));
}

constexpr uint32_t kDirectSumBlocks = 2048;  // <= this many blocks: every block sums its predecessors' aggregates directly

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_march_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
              const MarchParams p, const uint32_t N, const uint32_t M, const float* __restrict__ nears,
              const float* __restrict__ fars, const float* __restrict__ noises, float* __restrict__ xyzs,
              float* __restrict__ dirs, float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter,
              unsigned long long* __restrict__ scratch, const float* __restrict__ occ_box) {
    extern __shared__ float s_tl[];  // WARPS x max_steps visited t values
    __shared__ uint32_t s_cnt[WARPS];
    __shared__ uint32_t s_excl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n = blockIdx.x * WARPS + warp;
    float* tl = s_tl + (size_t)warp * p.max_steps;
    const Group<32> grp;

    Ray r;
    float t0 = 0.f, far = 0.f;
    const bool active = n < N;
    if (active) {
        r = make_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        const float near = nears[n];
        far = clip_far_to_box(occ_box, r, fars[n]);  // exact: nothing is sampled beyond the occupied cells (lnrf_march_rays_train_clipped)
        t0 = f_fma(f_clamp(f_mul(near, p.dt_gamma), p.dt_min, p.dt_max), noises[n], near);  // raymarching.cu:348-351
    } else {
        r = Ray{};
    }
    const uint32_t count = march_group<32, false>(grp, p, r, grid, t0, far, p.max_steps, active,
                                                  [&](uint32_t rank, float s, float, const Probe&, float) { tl[rank] = s; });
    if (lane == 0) s_cnt[warp] = count;
    __syncthreads();

    // ---- slot assignment: in-block prefix + prefix over blocks (status = flag<<32 | value) ----
    // scratch[0] = 1 + first sample row no ray wrote, scratch[1] = incoming counter[0], scratch[2..] = status words
    unsigned long long* status = scratch + 2;
    const uint32_t b = blockIdx.x;
    if (gridDim.x <= kDirectSumBlocks) {
        // Small grids (the 4096-ray training batch is 1024 blocks, all resident at once and all done marching at about the same
        // time): a look-back chain would resolve 32 blocks per L2 round trip, one hop after the other.  Instead every block
        // publishes its aggregate and adds up the aggregates of ALL its predecessors directly -- b loads spread over the block's
        // threads, no chain.  Integer sums: the result does not depend on the order.  Block 0 folds the incoming counter into its
        // aggregate, so nobody else reads counter[0] (the last block overwrites it).  Waiting only ever targets lower block ids.
        __shared__ uint32_t s_part[WARPS];
        if (threadIdx.x == 0) {
            uint32_t agg = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) agg += s_cnt[w];
            if (b == 0) {
                const uint32_t c0 = (uint32_t)counter[0];  // slots continue from the incoming counter, like the reference's atomicAdd
                scratch[1] = (unsigned long long)c0;
                if (c0 > M) scratch[0] = 1ull;  // nothing can be written at all: everything is zero-fill
                agg += c0;
            }
            st_relaxed_u64(status + b, (1ull << 32) | agg);
        }
        uint32_t part = 0;
        for (uint32_t j = threadIdx.x; j < b; j += WARPS * 32) {
            unsigned long long st;
            do {
                st = ld_relaxed_u64(status + j);
            } while ((st >> 32) == 0ull);
            part += (uint32_t)st;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
        if (lane == 0) s_part[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t excl = 0, agg = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) { excl += s_part[w]; agg += s_cnt[w]; }
            if (b == 0) excl = (uint32_t)scratch[1];  // block 0's own rays start at the incoming counter
            s_excl = excl;
            if (b == gridDim.x - 1) {
                const uint32_t total_end = excl + agg;
                counter[0] = (int)total_end;
                counter[1] += (int)N;
                if (total_end <= M) scratch[0] = (unsigned long long)total_end + 1ull;  // first row nobody wrote
            }
        }
    } else if (warp == 0) {
        // large grids: decoupled look-back (blocks complete in waves, inclusive prefixes are found within a window or two)
        uint32_t agg = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) agg += s_cnt[w];
        uint32_t excl = 0;
        if (b == 0) {
            excl = (uint32_t)counter[0];  // slots continue from the incoming counter, like the reference's atomicAdd
            if (lane == 0) {
                scratch[1] = (unsigned long long)excl;
                if (excl > M) scratch[0] = 1ull;  // nothing can be written at all: everything is zero-fill
            }
        } else {
            if (lane == 0) st_relaxed_u64(status + b, (1ull << 32) | agg);
            int idx = (int)b - 1;
            while (true) {
                const int j = idx - lane;
                unsigned long long st;
                do {
                    st = (j >= 0) ? ld_relaxed_u64(status + j) : (2ull << 32);
                } while (__any_sync(kFull, (st >> 32) == 0ull));
                const unsigned incl = __ballot_sync(kFull, (st >> 32) == 2ull);
                uint32_t v = (uint32_t)st;
                if (incl) {
                    const int first = __ffs(incl) - 1;
                    if (lane > first) v = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
                excl += v;
                if (incl) break;
                idx -= 32;
            }
        }
        if (lane == 0) {
            st_relaxed_u64(status + b, (2ull << 32) | (unsigned long long)(excl + agg));
            s_excl = excl;
            if (b == gridDim.x - 1) {
                const uint32_t total_end = excl + agg;
                counter[0] = (int)total_end;
                counter[1] += (int)N;
                if (total_end <= M) scratch[0] = (unsigned long long)total_end + 1ull;  // first row nobody wrote
            }
        }
    }
    __syncthreads();
    if (!active) return;
    uint32_t offset = s_excl;
    for (int w = 0; w < warp; w++) offset += s_cnt[w];

    if (lane == 0) {
        rays[(size_t)n * 3] = (int)n;
        rays[(size_t)n * 3 + 1] = (int)offset;
        rays[(size_t)n * 3 + 2] = (int)count;
    }
    if (count == 0) return;
    if (offset + count > M) {  // dropped (raymarching.cu:416); the first dropped ray marks where zero-fill starts
        if (offset <= M && lane == 0) scratch[0] = (unsigned long long)offset + 1ull;
        return;
    }
    __syncwarp();

    // ---- coalesced writes straight from the parked t values ----
    const float o3[3] = {r.ox, r.oy, r.oz};
    const float d3[3] = {r.dx, r.dy, r.dz};
    warp_write_flat(xyzs, (size_t)offset * 3, count * 3, lane, [&](uint32_t e) {
        const uint32_t k = e / 3u, c = e - 3u * k;
        const float oc = c == 0 ? o3[0] : (c == 1 ? o3[1] : o3[2]);
        const float dc = c == 0 ? d3[0] : (c == 1 ? d3[1] : d3[2]);
        return f_clamp(f_fma(tl[k], dc, oc), p.neg_bound, p.bound);
    });
    warp_write_flat(dirs, (size_t)offset * 3, count * 3, lane, [&](uint32_t e) {
        const uint32_t c = e % 3u;
        return c == 0 ? d3[0] : (c == 1 ? d3[1] : d3[2]);
    });
    float2* dl = reinterpret_cast<float2*>(deltas) + offset;
    for (uint32_t k = lane; k < count; k += 32) {
        const float s = tl[k];
        const float dt = p.dt_const ? p.dt0 : march_dt(p, s);
        const float after = f_add(s, dt);
        float last = t0;
        if (k > 0) {
            const float sp = tl[k - 1];
            last = f_add(sp, p.dt_const ? p.dt0 : march_dt(p, sp));
        }
        st_cs(dl + k, make_float2(dt, f_add(after, -last)));  // (dt, t - last_t), raymarching.cu:459-462
    }
}

// Zero rows [E, M) of the three sample buffers (what torch.zeros leaves behind in the reference wrapper,
// raymarching.py:205-207) and re-arm the look-back words for the next launch.
__global__ void __launch_bounds__(256)
k_march_train_tail(unsigned long long* __restrict__ scratch, uint32_t nblocks, float* __restrict__ xyzs,
                   float* __restrict__ dirs, float* __restrict__ deltas, uint32_t M) {
    const unsigned long long e1 = scratch[0];
    size_t head = (size_t)scratch[1];  // rows below the incoming counter value are nobody's: zero them too
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < nblocks; i += nth) scratch[2 + i] = 0ull;
    if (head > M) head = M;
    for (size_t i = tid; i < head * 3; i += nth) {
        xyzs[i] = 0.f;
        dirs[i] = 0.f;
    }
    for (size_t i = tid; i < head * 2; i += nth) deltas[i] = 0.f;
    if (e1 == 0ull) return;
    const size_t E = (size_t)(e1 - 1ull);
    if (E >= M) return;
    for (size_t i = E * 3 + tid; i < (size_t)M * 3; i += nth) {
        xyzs[i] = 0.f;
        dirs[i] = 0.f;
    }
    for (size_t i = E * 2 + tid; i < (size_t)M * 2; i += nth) deltas[i] = 0.f;
}

// =========================================================================================================
// training compositing (warp per ray)
// =========================================================================================================

// raymarching.cu:500-577, one warp per ray: returns the (warp-reduced) sums in every lane
struct RaySums {
    float r, g, b, ws, d;
};

// KEEP: plain loads (the fused forward+backward kernel reads the samples again right away) instead of streaming ones
template <bool KEEP = false>
__device__ __forceinline__ RaySums composite_ray_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                     const float* __restrict__ deltas, uint32_t offset, uint32_t num_steps,
                                                     uint32_t M, float T_thresh, int lane) {
    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    if (num_steps != 0 && offset + num_steps <= M) {
        const float* ps = sigmas + offset;
        const float* pc = rgbs + (size_t)offset * 3;
        const float2* pl = reinterpret_cast<const float2*>(deltas) + offset;
        float T_carry = 1.0f, t_carry = 0.0f;
        // the loads of chunk k + 1 are issued before the scans of chunk k: a ray's life is a chain of dependent round trips, and at
        // large batch sizes the bytes in flight per SM are what bounds these kernels (one warp = one ray = 768 B per chunk)
        struct Chunk { float sg, dx, dy, cr, cg, cb; };
        auto fetch = [&](uint32_t base) {
            Chunk c{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const uint32_t k = base + lane;
            if (k < num_steps) {
                const float2 dl = KEEP ? __ldg(pl + k) : __ldcs(pl + k);
                c.sg = KEEP ? __ldg(ps + k) : __ldcs(ps + k);
                c.dx = dl.x; c.dy = dl.y;
                if (KEEP) {
                    c.cr = __ldg(pc + (size_t)k * 3); c.cg = __ldg(pc + (size_t)k * 3 + 1); c.cb = __ldg(pc + (size_t)k * 3 + 2);
                } else {
                    c.cr = __ldcs(pc + (size_t)k * 3); c.cg = __ldcs(pc + (size_t)k * 3 + 1); c.cb = __ldcs(pc + (size_t)k * 3 + 2);
                }
            }
            return c;
        };
        Chunk cur = fetch(0);
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t k = base + lane;
            const bool act = k < num_steps;
            Chunk nxt{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (base + 32 < num_steps) nxt = fetch(base + 32);
            const float alpha = act ? 1.0f - __expf(-cur.sg * cur.dx) : 0.f;
            const float cr = cur.cr, cg = cur.cg, cb = cur.cb, dd = cur.dy;
            cur = nxt;
            float P = 1.0f - alpha;  // inclusive product scan of (1 - alpha)
            float S = dd;            // inclusive sum scan of the depth deltas
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float Pn = __shfl_up_sync(kFull, P, o), Sn = __shfl_up_sync(kFull, S, o);
                if (lane >= o) { P *= Pn; S += Sn; }
            }
            float Pex = __shfl_up_sync(kFull, P, 1);
            if (lane == 0) Pex = 1.0f;
            const float T_before = T_carry * Pex, T_after = T_carry * P;
            // the reference accumulates the sample that drives T below the threshold, then stops (:554-557)
            const unsigned stop = __ballot_sync(kFull, act && (T_after < T_thresh));
            const bool use = act && (stop == 0u || lane <= (__ffs(stop) - 1));
            if (use) {
                const float w = alpha * T_before;
                r += w * cr; g += w * cg; b += w * cb;
                d += w * (t_carry + S);
                ws += w;
            }
            if (stop) break;
            T_carry = __shfl_sync(kFull, T_after, 31);
            t_carry += __shfl_sync(kFull, S, 31);
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
    return RaySums{r, g, b, ws, d};
}

__global__ void __launch_bounds__(256)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                      const int* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                      float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    // (the loop runs once with the launch below: one block per 8 rays)
    const int lane = threadIdx.x & 31;
    for (uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += gridDim.x * 8u) {
        const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3), offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1),
                       num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);
        const RaySums a = composite_ray_fwd(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane);
        if (lane == 0) {
            weights_sum[index] = a.ws;
            depth[index] = a.d;
            image[(size_t)index * 3] = a.r;
            image[(size_t)index * 3 + 1] = a.g;
            image[(size_t)index * 3 + 2] = a.b;
        }
    }
}

// Row f-5: the tail of the training step after the network -- compositing, background blend
// (renderer.py:326), depth normalisation (:328) and the trainer's MSE (nerf/utils.py:592 `criterion(pred, gt).mean(-1)`,
// :633 `.mean()`) -- as ONE launch.  The reference spends ~12 elementwise / reduction launches here (and as many in the
// backward).  The loss is reduced deterministically: per-block partials, then the last block out adds them in index order.
struct LossArgs {
    const float* gt;         // [N,3] target colours
    const float* bg;         // [N,3] per-pixel background or null -> bg_scalar (bg_color = 1 in the trainer's default)
    float bg_scalar;
    const float* nears;      // null: depth is left unscaled
    const float* fars;
    float* image_raw;        // [N,3] composite before the blend, saved for the backward
    float* partial;          // [gridDim.x]
    unsigned int* ticket;    // zero before first use; the kernel re-arms it
    float* loss;             // [1]
    const float* grad_loss;  // backward only: dL_total/dloss on the device (the AMP loss scale)
};

// per-ray part of the loss forward: blended image, depth, squared error (returned in every lane; lane 0 stores)
template <bool KEEP>
__device__ __forceinline__ float composite_loss_ray_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                        const float* __restrict__ deltas, uint32_t index, uint32_t offset,
                                                        uint32_t num_steps, uint32_t M, float T_thresh, const LossArgs& la,
                                                        float* __restrict__ weights_sum, float* __restrict__ depth,
                                                        float* __restrict__ image, int lane, RaySums& a, float v[3]) {
    a = composite_ray_fwd<KEEP>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane);
    const size_t i3 = (size_t)index * 3;
    const float raw[3] = {a.r, a.g, a.b};
    const float omw = __fadd_rn(1.0f, -a.ws);
    float e = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float bgc = la.bg ? __ldg(la.bg + i3 + c) : la.bg_scalar;
        v[c] = __fadd_rn(raw[c], __fmul_rn(omw, bgc));  // image + (1 - weights_sum) * bg_color
        const float df = __fadd_rn(v[c], -__ldg(la.gt + i3 + c));
        e = __fadd_rn(e, __fmul_rn(df, df));
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            la.image_raw[i3 + c] = raw[c];
            image[i3 + c] = v[c];
        }
        weights_sum[index] = a.ws;
        float dpt = a.d;
        if (la.nears) {  // clamp(depth - nears, min=0) / (fars - nears)
            const float nr = __ldg(la.nears + index);
            dpt = __fdiv_rn(fmaxf(__fadd_rn(dpt, -nr), 0.0f), __fadd_rn(__ldg(la.fars + index), -nr));
        }
        depth[index] = dpt;
    }
    return e;
}

// deterministic loss reduction: per-block partials, then the last block out adds them in index order
__device__ __forceinline__ void composite_loss_reduce(float err, const LossArgs& la, uint32_t N, int lane, int warp) {
    __shared__ float s_err[8];
    __shared__ bool s_last;
    if (lane == 0) s_err[warp] = err;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += s_err[w];
        la.partial[blockIdx.x] = s;
        __threadfence();
        s_last = atomicAdd(la.ticket, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last) {  // the last block out: fixed-order sum of the partials (thread-strided, then warps, then the 8 warp sums)
        __threadfence();
        float s = 0.f;
        uint32_t i = threadIdx.x;
        for (; i + 7u * 256u < gridDim.x; i += 8u * 256u) {  // the order of one load per trip, eight of them in flight (large batches)
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = __ldcg(la.partial + i + (uint32_t)u * 256u);
#pragma unroll
            for (int u = 0; u < 8; u++) s += x[u];
        }
        for (; i < gridDim.x; i += 256) s += __ldcg(la.partial + i);
        s = warp_sum(s);
        __syncthreads();  // s_err is being reused
        if (lane == 0) s_err[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += s_err[w];
            la.loss[0] = tot / (3.0f * (float)N);
            *la.ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(256)
k_composite_loss_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                     const int* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh, const LossArgs la,
                     float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    pdl_trigger();
    pdl_wait();  // nothing below may run before the kernels ahead on the stream are complete (common.cuh)
    // the grid is capped (composite_loss_blocks): beyond one wave a warp takes several rays, so that the block-level ticket of the
    // loss reduction -- a fence and an atomic round trip during which the whole block holds its registers -- is paid once per block,
    // not once per 8 rays (65 536 rays: 62 -> see profiles/ large-batch table)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float err = 0.f;
    for (uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += gridDim.x * 8u) {
        const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3), offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1),
                       num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);
        RaySums a;
        float v[3];
        err += composite_loss_ray_fwd<false>(sigmas, rgbs, deltas, index, offset, num_steps, M, T_thresh, la, weights_sum, depth, image, lane, a, v);
    }
    composite_loss_reduce(err, la, N, lane, warp);
}

__device__ __forceinline__ void warp_zero_range(float* __restrict__ p, size_t lo, size_t hi, int lane) {
    for (size_t i = lo + lane; i < hi; i += 32) p[i] = 0.f;
}

// the gradient rows no fitting ray covers (canonical layout: [first dropped ray's offset, M), or [end of last ray, M), and rows below
// the first ray's offset): zero-filled by the warps of the rays next to them
__device__ __forceinline__ void composite_zero_uncovered(uint32_t n, uint32_t N, uint32_t offset, uint32_t num_steps, bool fits, uint32_t M,
                                                         float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, int lane) {
    size_t z0 = M;
    if (num_steps != 0 && !fits && offset <= M) z0 = offset;
    else if (n == N - 1 && fits) z0 = (size_t)offset + num_steps;
    if (z0 < M) {
        warp_zero_range(grad_sigmas, z0, M, lane);
        warp_zero_range(grad_rgbs, z0 * 3, (size_t)M * 3, lane);
    }
    if (n == 0 && offset > 0) {
        const size_t z1 = offset < M ? offset : M;
        warp_zero_range(grad_sigmas, 0, z1, lane);
        warp_zero_range(grad_rgbs, 0, z1 * 3, lane);
    }
}

// raymarching.cu:601-682 for one ray (one warp): gradients of its samples from dL/dimage (gr, gg, gb), dL/dweights_sum (gws) and the
// ray's forward results
template <bool ZERO_FILL>
__device__ __forceinline__ void composite_ray_bwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                  const float* __restrict__ deltas, uint32_t offset, uint32_t num_steps, float T_thresh,
                                                  float gws, float gr, float gg, float gb, float r_final, float g_final, float b_final,
                                                  float ws_final, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, int lane) {
    const float* ps = sigmas + offset;
    const float* pc = rgbs + (size_t)offset * 3;
    const float2* pl = reinterpret_cast<const float2*>(deltas) + offset;
    float* gs = grad_sigmas + offset;
    float* gc = grad_rgbs + (size_t)offset * 3;
    const float ws_term = gws * (1.0f - ws_final);

    float T_carry = 1.0f, r_carry = 0.f, g_carry = 0.f, b_carry = 0.f;
    uint32_t base = 0;
    struct Chunk { float sg, dt, cr, cg, cb; };  // chunk k + 1 is requested before the scans of chunk k (see composite_ray_fwd)
    auto fetch = [&](uint32_t at) {
        Chunk c{0.f, 0.f, 0.f, 0.f, 0.f};
        const uint32_t k = at + lane;
        if (k < num_steps) {
            c.dt = __ldcs(pl + k).x;
            c.sg = __ldcs(ps + k);
            c.cr = __ldcs(pc + (size_t)k * 3); c.cg = __ldcs(pc + (size_t)k * 3 + 1); c.cb = __ldcs(pc + (size_t)k * 3 + 2);
        }
        return c;
    };
    Chunk cur = fetch(0);
    for (; base < num_steps; base += 32) {
        const uint32_t k = base + lane;
        const bool act = k < num_steps;
        Chunk nxt{0.f, 0.f, 0.f, 0.f, 0.f};
        if (base + 32 < num_steps) nxt = fetch(base + 32);
        const float dt = cur.dt, cr = cur.cr, cg = cur.cg, cb = cur.cb;
        const float alpha = act ? 1.0f - __expf(-cur.sg * dt) : 0.f;
        cur = nxt;
        float P = 1.0f - alpha;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float Pn = __shfl_up_sync(kFull, P, o);
            if (lane >= o) P *= Pn;
        }
        float Pex = __shfl_up_sync(kFull, P, 1);
        if (lane == 0) Pex = 1.0f;
        const float T_before = T_carry * Pex, T_after = T_carry * P;
        const float w = alpha * T_before;
        float ar = w * cr, ag = w * cg, ab = w * cb;  // inclusive sums of the accumulated colour
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float xr = __shfl_up_sync(kFull, ar, o), xg = __shfl_up_sync(kFull, ag, o), xb = __shfl_up_sync(kFull, ab, o);
            if (lane >= o) { ar += xr; ag += xg; ab += xb; }
        }
        ar += r_carry; ag += g_carry; ab += b_carry;
        const unsigned stop = __ballot_sync(kFull, act && (T_after < T_thresh));
        const bool use = act && (stop == 0u || lane <= (__ffs(stop) - 1));
        if (use) {
            __stcs(gc + (size_t)k * 3, gr * w);
            __stcs(gc + (size_t)k * 3 + 1, gg * w);
            __stcs(gc + (size_t)k * 3 + 2, gb * w);
            __stcs(gs + k, dt * (gr * (T_after * cr - (r_final - ar)) + gg * (T_after * cg - (g_final - ag)) +
                                 gb * (T_after * cb - (b_final - ab)) + ws_term));
        } else if (ZERO_FILL && act) {
            __stcs(gc + (size_t)k * 3, 0.f); __stcs(gc + (size_t)k * 3 + 1, 0.f); __stcs(gc + (size_t)k * 3 + 2, 0.f);
            __stcs(gs + k, 0.f);
        }
        if (stop) { base += 32; break; }
        T_carry = __shfl_sync(kFull, T_after, 31);
        r_carry = __shfl_sync(kFull, ar, 31);
        g_carry = __shfl_sync(kFull, ag, 31);
        b_carry = __shfl_sync(kFull, ab, 31);
    }
    if (ZERO_FILL && base < num_steps) {  // samples behind the early-out keep zero gradients
        warp_zero_range(gs, base, num_steps, lane);
        warp_zero_range(gc, (size_t)base * 3, (size_t)num_steps * 3, lane);
    }
}

// raymarching.cu:601-682
// LOSS (row f-5): dL/dimage and dL/dweights_sum are not read but formed here from the blended image, the targets and the
// device-side scalar dL/dloss: g = dloss * 2 (image - gt) / (3 N), gws = -sum_c g_c bg_c; `image` is then the blended image and
// la.image_raw the composite the C_final - C_acc term needs.
template <bool ZERO_FILL, bool LOSS>
__global__ void __launch_bounds__(256)
k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                      const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                      const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ image,
                      uint32_t M, uint32_t N, float T_thresh, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs,
                      const LossArgs la) {
    pdl_trigger();
    pdl_wait();  // nothing below may run before the kernels ahead on the stream are complete (common.cuh)
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3), offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1),
                   num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);
    const bool fits = offset + num_steps <= M;
    if (ZERO_FILL) composite_zero_uncovered(n, N, offset, num_steps, fits, M, grad_sigmas, grad_rgbs, lane);
    if (num_steps == 0 || !fits) return;

    float gws, gr, gg, gb, r_final, g_final, b_final;
    const size_t i3 = (size_t)index * 3;
    if (LOSS) {
        const float gsc = __ldg(la.grad_loss) * (2.0f / (3.0f * (float)N));
        gr = gsc * (image[i3] - __ldg(la.gt + i3));
        gg = gsc * (image[i3 + 1] - __ldg(la.gt + i3 + 1));
        gb = gsc * (image[i3 + 2] - __ldg(la.gt + i3 + 2));
        gws = la.bg ? -(gr * __ldg(la.bg + i3) + gg * __ldg(la.bg + i3 + 1) + gb * __ldg(la.bg + i3 + 2))
                    : -(gr + gg + gb) * la.bg_scalar;
        r_final = la.image_raw[i3]; g_final = la.image_raw[i3 + 1]; b_final = la.image_raw[i3 + 2];
    } else {
        gws = grad_weights_sum[index];
        gr = grad_image[i3]; gg = grad_image[i3 + 1]; gb = grad_image[i3 + 2];
        r_final = image[i3]; g_final = image[i3 + 1]; b_final = image[i3 + 2];
    }
    composite_ray_bwd<ZERO_FILL>(sigmas, rgbs, deltas, offset, num_steps, T_thresh, gws, gr, gg, gb, r_final, g_final, b_final,
                                 weights_sum[index], grad_sigmas, grad_rgbs, lane);
}

// Forward AND backward of the tail in one launch: the loss's gradient with respect to itself is a device scalar known before the
// forward runs (the AMP loss scale), so the warp that composited a ray turns straight around and writes the ray's sample gradients
// while its samples are still in L1.  Same expressions as k_composite_loss_fwd followed by k_composite_train_bwd<true, true>: the
// results are bit-identical to the two launches.
__global__ void __launch_bounds__(256)
k_composite_loss_fwd_bwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                         const int* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh, const LossArgs la,
                         float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
                         float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    pdl_trigger();
    pdl_wait();  // nothing below may run before the kernels ahead on the stream are complete (common.cuh)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float err = 0.f;
    for (uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += gridDim.x * 8u) {
        const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3), offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1),
                       num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);
        RaySums a;
        float v[3];
        err += composite_loss_ray_fwd<true>(sigmas, rgbs, deltas, index, offset, num_steps, M, T_thresh, la, weights_sum, depth, image, lane, a, v);
        const bool fits = offset + num_steps <= M;
        composite_zero_uncovered(n, N, offset, num_steps, fits, M, grad_sigmas, grad_rgbs, lane);
        if (num_steps != 0 && fits) {
            const size_t i3 = (size_t)index * 3;
            const float gsc = __ldg(la.grad_loss) * (2.0f / (3.0f * (float)N));
            const float gr = gsc * (v[0] - __ldg(la.gt + i3)), gg = gsc * (v[1] - __ldg(la.gt + i3 + 1)), gb = gsc * (v[2] - __ldg(la.gt + i3 + 2));
            const float gws = la.bg ? -(gr * __ldg(la.bg + i3) + gg * __ldg(la.bg + i3 + 1) + gb * __ldg(la.bg + i3 + 2))
                                    : -(gr + gg + gb) * la.bg_scalar;
            composite_ray_bwd<true>(sigmas, rgbs, deltas, offset, num_steps, T_thresh, gws, gr, gg, gb, a.r, a.g, a.b, a.ws, grad_sigmas,
                                    grad_rgbs, lane);
        }
    }
    composite_loss_reduce(err, la, N, lane, warp);
}

// =========================================================================================================
// inference / distillation march (group of 8 lanes per alive ray, fixed n_step slots per ray)
// =========================================================================================================

// Device-side round control (row f-3 of SURVEY.md section 8): the inference loop of NeRFRenderer.run_cuda
// (nerf/renderer.py:353-379) sizes every round on the host from `n_alive` -- one device->host synchronisation per round,
// 100+ rounds per frame.  With a control block in device memory the kernels read the round's geometry themselves, the
// compaction kernel computes the next round's, and the host only looks at it every few rounds.
//   ctl[0] n_alive   ctl[1] n_step = clamp(n_rays / n_alive, 1, 8)   ctl[2] steps marched so far (renderer.py: `step`)
//   ctl[3] rows = n_alive * n_step rounded up PAST the next multiple of 128 (raymarching.py:318-320)
//   ctl[4] n_rays    ctl[5] max_steps    ctl[6] finished (n_alive == 0 or step >= max_steps)    ctl[7] rounds executed
//   ctl[8] survivors of the running compaction (internal)    ctl[9] sample slots marched so far
//   ctl[10] row budget of the rounds after the first (n_rays: the reference schedule; more: see lnrf_render_desc.sample_rows)
//   ctl[11] most samples a ray takes per round after the first (8: the reference schedule)
//   ctl[12] schedule-dependence flag: set when the frame's result COULD depend on where the round boundaries fall (see below)
//   ctl[13] rays that raised ctl[12]   ctl[14] the max_steps cap cut rays off (the frame goes to the reference schedule)
//   ctl[15] rows of the running COMPACT round (k_march_infer_compact; 0 on slot-layout rounds and when the frame is finished)
//   ctl[16] length of a prescribed n_step sequence (0: none)   ctl[17,18] its device address   ctl[19,20] address of the per-ray
//   completed-step counters   ctl[21,22] address of the per-ray schedule-dependence flags (see lnrf_render_desc)
enum { kCtlAlive = 0, kCtlStep = 1, kCtlSteps = 2, kCtlRows = 3, kCtlRays = 4, kCtlMaxSteps = 5, kCtlFinished = 6, kCtlRounds = 7,
       kCtlBudget = 10, kCtlStepCap = 11, kCtlInexact = 12, kCtlSeqLen = 16, kCtlSeqPtr = 17, kCtlStepsPtr = 19, kCtlFlagsPtr = 21 };
template <typename T>
__device__ __forceinline__ T* ctl_ptr(const int* ctl, int slot) {
    return reinterpret_cast<T*>(((unsigned long long)(unsigned int)ctl[slot + 1] << 32) | (unsigned long long)(unsigned int)ctl[slot]);
}
__device__ __forceinline__ void ctl_set_ptr(int* ctl, int slot, const void* p) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(p);
    ctl[slot] = (int)(unsigned int)(a & 0xffffffffull);
    ctl[slot + 1] = (int)(unsigned int)(a >> 32);
}
// Why a schedule other than the reference's n_step rule can be bit-identical to it.  Between rounds a ray's t travels through
// rays_t, which composite_rays rebuilds as rays_t + sum of deltas[.][1] (raymarching.cu:1006), each delta being fl(t_after -
// t_prev) from the marcher (:789).  When t_after / t_prev <= 2 that difference is exact (Sterbenz) and fl(t_prev + delta) is
// t_after itself, so the rebuilt rays_t equals the marcher's own t wherever the round ends: boundaries leave no trace.  Only an
// inexact delta (a skip longer than t itself: cameras inside the volume, min_near << gap) injects a rounding error whose position
// depends on the schedule.  The marcher checks fl(t_prev + delta) == t_after for every sample it emits and raises ctl[12]
// otherwise; so does a frame that runs into the max_steps cap with rays still alive.  No flag = every schedule gives the
// reference's bits (measured: the lego-shape view, 91 vs 16 rounds, image / depth / weights identical).

__device__ __forceinline__ void ctl_set_round(int* ctl, uint32_t n_alive) {
    const uint32_t n_rays = (uint32_t)ctl[kCtlRays];
    bool fin = n_alive == 0u || (uint32_t)ctl[kCtlSteps] >= (uint32_t)ctl[kCtlMaxSteps];
    const uint32_t budget = ctl[kCtlRounds] > 0 ? (uint32_t)ctl[kCtlBudget] : n_rays;
    uint32_t n_step = n_alive ? budget / n_alive : 1u;
    const uint32_t step_cap = ctl[kCtlRounds] > 0 ? (uint32_t)ctl[kCtlStepCap] : 8u;
    n_step = n_step > step_cap ? step_cap : (n_step < 1u ? 1u : n_step);
    if (ctl[kCtlSeqLen] > 0) {  // prescribed schedule: round r takes seq[r] samples per ray, whatever n_alive is
        const uint32_t r = (uint32_t)ctl[kCtlRounds];
        if (r >= (uint32_t)ctl[kCtlSeqLen]) fin = true;
        else n_step = (uint32_t)ctl_ptr<const int>(ctl, kCtlSeqPtr)[r];
    }
    uint32_t rows = n_alive * n_step;
    rows += 128u - rows % 128u;
    ctl[15] = 0;  // kCtlCompactRows: the compact marcher of the round publishes it; 0 until then and when the frame is finished
    ctl[kCtlAlive] = fin ? 0 : (int)n_alive;
    ctl[kCtlStep] = (int)n_step;
    ctl[kCtlRows] = fin ? 0 : (int)rows;
    ctl[kCtlFinished] = fin ? 1 : 0;
}

// group of G lanes per alive ray (8 covers the <= 8 samples per ray of a round; see march_infer_dev_launch for the first round)
constexpr int kInferGroup = 8;

template <bool DISTILL, int G>
__global__ void __launch_bounds__(256)
k_march_infer(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
              const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
              const MarchParams p, const uint8_t* __restrict__ grid, const uint8_t* __restrict__ edit_grid,
              const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
              float* __restrict__ deltas, uint8_t* __restrict__ edit_occ, const float* __restrict__ noises,
              uint32_t M_rows, uint32_t n_groups, const int* __restrict__ ctl, const int g_lo, const int g_hi,
              const float* __restrict__ occ_box = nullptr) {
    const Group<G> grp;
    if (ctl) {  // device-driven round: geometry from the control block; this instantiation serves n_step in [g_lo, g_hi]
        n_alive = (uint32_t)ctl[kCtlAlive];
        n_step = (uint32_t)ctl[kCtlStep];
        M_rows = (uint32_t)ctl[kCtlRows];
        if (M_rows == 0u || (int)n_step < g_lo || (int)n_step > g_hi) return;
        n_groups = div_up(M_rows, n_step);
    }
    const uint32_t groups_per_pass = gridDim.x * (blockDim.x / G);
    const uint32_t passes = div_up(n_groups, groups_per_pass);
    for (uint32_t pass = 0; pass < passes; pass++) {  // uniform trip count: march_group uses full-warp votes
        const uint32_t g = pass * groups_per_pass + (blockIdx.x * blockDim.x + threadIdx.x) / G;
        const bool active = g < n_alive;
        Ray r = Ray{};
        float t = 0.f, far = 0.f;
        int index = 0;
        if (active) {
            index = __ldg(rays_alive + g);
            r = make_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
            far = clip_far_to_box(occ_box, r, fars[index]);
            const float t_in = rays_t[index];
            const float noise = noises ? noises[g] : 0.0f;
            t = f_fma(f_clamp(f_mul(t_in, p.dt_gamma), p.dt_min, p.dt_max), noise, t_in);  // raymarching.cu:746
        }
        const size_t row0 = (size_t)g * n_step;
        bool inexact = false;
        const uint32_t cnt = march_group<G, true>(
            grp, p, r, grid, t, far, n_step, active, [&](uint32_t rank, float s, float dt, const Probe& q, float prev_after) {
                const size_t row = row0 + rank;
                xyzs[row * 3] = q.x; xyzs[row * 3 + 1] = q.y; xyzs[row * 3 + 2] = q.z;
                dirs[row * 3] = r.dx; dirs[row * 3 + 1] = r.dy; dirs[row * 3 + 2] = r.dz;
                const float t_after = f_add(s, dt), d1 = f_add(t_after, -prev_after);
                reinterpret_cast<float2*>(deltas)[row] = make_float2(dt, d1);
                inexact |= f_add(prev_after, d1) != t_after;
                if (DISTILL) edit_occ[row] = (uint8_t)((__ldg(edit_grid + (q.index >> 3)) >> (q.index & 7u)) & 1u);
            });
        // Round 0 is exempt: it takes ONE sample per ray on every schedule (ctl_set_round), so the boundary after a ray's first sample
        // -- typically the one long skip from `near` to the first occupied cell, where t can more than double -- sits in the same
        // place whatever happens afterwards.
        inexact = grp.ballot(inexact) != 0u;     // the lanes of a group hold different samples of the ray: any of them raises it
        if (ctl && inexact && ctl[kCtlRounds] > 0) {
            int* c = const_cast<int*>(ctl);
            c[kCtlInexact] = 1;                  // benign race: every writer stores the same value
            if (grp.gl == 0) {
                atomicAdd(c + 13, 1);            // how many rays raised it
                uint8_t* flags = ctl_ptr<uint8_t>(ctl, kCtlFlagsPtr);
                if (flags) flags[index] = 1;
            }
        }
        if (g >= n_groups) continue;
        // zero-fill the slots this ray did not use, and whole padding groups (torch.zeros in raymarching.py:334-336)
        const size_t zlo = row0 + (active ? cnt : 0u);
        size_t zhi = row0 + n_step;
        if (zhi > M_rows) zhi = M_rows;
        for (size_t i = zlo * 3 + grp.gl; i < zhi * 3; i += G) { xyzs[i] = 0.f; dirs[i] = 0.f; }
        for (size_t i = zlo * 2 + grp.gl; i < zhi * 2; i += G) deltas[i] = 0.f;
        if (DISTILL)
            for (size_t i = zlo + grp.gl; i < zhi; i += G) edit_occ[i] = 0;
    }
}

// raymarching.cu:948-1035 and :1037-1142.  Thread per alive ray: at most n_step <= 8 samples, and the reference's
// sequential float order (T = 1 - weight_sum; t += delta) is kept so that rays_t and the kill pattern are exact.
template <bool DISTILL>
__device__ __forceinline__ void composite_infer_ray(const uint32_t n, const uint32_t n_step, const float T_thresh, int* __restrict__ rays_alive,
                  float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                  const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ weights_edit_sum,
                  float* __restrict__ depth, float* __restrict__ depth_edit, const uint8_t* __restrict__ edit_occ,
                  float* __restrict__ image, int* __restrict__ ray_steps = nullptr);

template <bool DISTILL>
__global__ void __launch_bounds__(256)
k_composite_infer(const uint32_t n_alive, const uint32_t n_step, const float T_thresh, int* __restrict__ rays_alive,
                  float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                  const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ weights_edit_sum,
                  float* __restrict__ depth, float* __restrict__ depth_edit, const uint8_t* __restrict__ edit_occ,
                  float* __restrict__ image, const int* __restrict__ ctl) {
    uint32_t na = n_alive, ns = n_step;
    int* ray_steps = nullptr;
    if (ctl) { na = (uint32_t)ctl[kCtlAlive]; ns = (uint32_t)ctl[kCtlStep]; ray_steps = ctl_ptr<int>(ctl, kCtlStepsPtr); }
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < na; n += gridDim.x * blockDim.x)
        composite_infer_ray<DISTILL>(n, ns, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, weights_edit_sum, depth,
                                     depth_edit, edit_occ, image, ray_steps);
}

template <bool DISTILL>
__device__ __forceinline__ void composite_infer_ray(const uint32_t n, const uint32_t n_step, const float T_thresh, int* __restrict__ rays_alive,
                  float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                  const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ weights_edit_sum,
                  float* __restrict__ depth, float* __restrict__ depth_edit, const uint8_t* __restrict__ edit_occ,
                  float* __restrict__ image, int* __restrict__ ray_steps) {
    const int index = rays_alive[n];
    const float* ps = sigmas + (size_t)n * n_step;
    const float* pc = rgbs + (size_t)n * n_step * 3;
    const float2* pl = reinterpret_cast<const float2*>(deltas) + (size_t)n * n_step;
    const uint8_t* pe = DISTILL ? edit_occ + (size_t)n * n_step : nullptr;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float weight_edit_sum = 0.f, d_edit = 0.f;
    if (DISTILL) { weight_edit_sum = weights_edit_sum[index]; d_edit = depth_edit[index]; }
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const float2 dl = __ldg(pl + step);
        if (dl.x == 0.f) break;
        const float alpha = 1.0f - __expf(-__ldg(ps + step) * dl.x);
        const float T = f_add(1.0f, -weight_sum);
        const float weight = f_mul(alpha, T);
        weight_sum = f_add(weight_sum, weight);
        if (DISTILL && pe[step]) {
            weight_edit_sum = f_add(weight_edit_sum, weight);
            d_edit = f_fma(weight, t, d_edit);  // t BEFORE the increment (:1098-1101); nvcc contracts w*t + d
        }
        t = f_add(t, dl.y);
        d = f_fma(weight, t, d);
        r = f_fma(weight, __ldg(pc + step * 3), r);
        g = f_fma(weight, __ldg(pc + step * 3 + 1), g);
        b = f_fma(weight, __ldg(pc + step * 3 + 2), b);
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    if (ray_steps) ray_steps[index] += (int)step;  // completed samples: the total is where the ray dies, on any schedule
    if (DISTILL) { weights_edit_sum[index] = weight_edit_sum; depth_edit[index] = d_edit; }
    weights_sum[index] = weight_sum;
    depth[index] = d;
    image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
}

// =========================================================================================================
// device-side alive-ray compaction (row f-3): stable, one launch, decoupled look-back over 1024-ray blocks
// =========================================================================================================
// scratch[0] = finished-block counter, scratch[1 + b] = look-back word of block b (flag << 32 | count).  The block
// that finishes last re-zeroes everything, so the scratch is ready for the next launch on the same stream.
__global__ void __launch_bounds__(1024)
k_compact_alive(const int* __restrict__ rays_alive, uint32_t n_alive, int* __restrict__ out, int* __restrict__ n_out,
                unsigned long long* __restrict__ scratch, int* __restrict__ ctl) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ctl) n_alive = (uint32_t)ctl[kCtlAlive];  // device-driven round: the grid covers the ray capacity
    // blocks past the last alive entry hold nothing and are not part of anyone's look-back: they only check out below
    const uint32_t active_blocks = n_alive ? div_up(n_alive, 1024u) : 1u;
    const bool idle_block = blockIdx.x >= active_blocks;
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    const int v = i < n_alive ? rays_alive[i] : -1;
    const bool keep = v >= 0;
    const unsigned bal = __ballot_sync(kFull, keep);
    if (lane == 0) s_warp[warp] = (uint32_t)__popc(bal);
    __syncthreads();
    unsigned long long* status = scratch + 1;
    if (warp == 0 && !idle_block) {
        uint32_t c = s_warp[lane], incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += n;
        }
        s_warp[lane] = incl - c;  // exclusive prefix of the warps inside the block
        const uint32_t agg = __shfl_sync(kFull, incl, 31);
        uint32_t excl = 0;
        const uint32_t b = blockIdx.x;
        if (b > 0) {
            if (lane == 0) st_relaxed_u64(status + b, (1ull << 32) | agg);
            int idx = (int)b - 1;
            while (true) {
                const int j = idx - lane;
                unsigned long long st;
                do {
                    st = (j >= 0) ? ld_relaxed_u64(status + j) : (2ull << 32);
                } while (__any_sync(kFull, (st >> 32) == 0ull));
                const unsigned inc = __ballot_sync(kFull, (st >> 32) == 2ull);
                uint32_t x = (uint32_t)st;
                if (inc) {
                    const int first = __ffs(inc) - 1;
                    if (lane > first) x = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
                excl += x;
                if (inc) break;
                idx -= 32;
            }
        }
        if (lane == 0) {
            st_relaxed_u64(status + b, (2ull << 32) | (unsigned long long)(excl + agg));
            s_excl = excl;
            if (b == active_blocks - 1) {
                if (n_out) *n_out = (int)(excl + agg);
                if (ctl) ctl[kCtlRounds + 1] = (int)(excl + agg);  // parked; the last block out publishes the next round
            }
        }
    }
    __syncthreads();
    if (keep && !idle_block) out[s_excl + s_warp[warp] + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = v;
    // last block out cleans up
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(scratch, 1ull) == (unsigned long long)gridDim.x - 1ull;
    }
    __syncthreads();
    if (s_last) {
        for (uint32_t k = threadIdx.x; k < gridDim.x + 1u; k += 1024u) scratch[k] = 0ull;
        if (ctl && threadIdx.x == 0) {  // every block has read ctl[kCtlAlive] and written its survivors: set up the next round
            __threadfence();
            if (!ctl[kCtlFinished]) {
                ctl[kCtlSteps] += ctl[kCtlStep];
                ctl[kCtlRounds] += 1;
                ctl[kCtlRounds + 2] += ctl[15] > 0 ? ctl[15] : ctl[kCtlRows];  // sample rows marched so far (compact rounds: the real ones)
                const uint32_t survivors = (uint32_t)((volatile int*)ctl)[kCtlRounds + 1];
                if (survivors > 0u && (uint32_t)ctl[kCtlSteps] >= (uint32_t)ctl[kCtlMaxSteps]) {  // the cap cut rays off
                    ctl[kCtlInexact] = 1;
                    ctl[14] = 1;
                }
                ctl_set_round(ctl, survivors);
            }
        }
    }
}

// =========================================================================================================
// compact rounds of the fast schedule (rounds >= 1): a ray's samples sit behind the previous ray's, no empty slots
// =========================================================================================================
// The reference's layout gives every alive ray n_step slots (raymarching.cu:760) and zero-fills what the ray does not use; with up
// to 64 samples per ray per round that is ~15 % of the rows of a frame going through encoder and network as zeros.  Here the
// marcher parks a ray's sample positions in shared memory, the groups of a block and then the blocks (decoupled look-back, as in
// k_march_train) agree on offsets, and the samples are written back to back; (offset, count) per alive ray is what the compositor
// reads.  Same positions, same deltas, same per-ray arithmetic: only the row a sample lives in changes, which no output depends on.
//   status words (scratch): flag << 32 | count, zero before the launch (k_composite_infer_compact re-arms them)
//   ctl[15] = rows of this round, rounded up to a multiple of 128 (the pad rows are zero-filled here)
constexpr int kCtlCompactRows = 15;
constexpr int kCompactCap = 64;  // most samples a ray takes per round (render_begin: step_cap <= 64)

constexpr int kCompactG = 8;          // lanes per ray (default)
constexpr int kCompactThreads = 128;  // 16 rays per block: the block waits for its longest march before the offsets are known

template <bool DISTILL, int G>
__global__ void __launch_bounds__(kCompactThreads)
k_march_infer_compact(const int* __restrict__ rays_alive, const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                      const float* __restrict__ rays_d, const MarchParams p, const uint8_t* __restrict__ grid,
                      const uint8_t* __restrict__ edit_grid, const float* __restrict__ fars, float* __restrict__ xyzs,
                      float* __restrict__ dirs, float* __restrict__ deltas, uint8_t* __restrict__ edit_occ, int* __restrict__ ray_off,
                      int* __restrict__ ray_cnt, unsigned long long* __restrict__ status, int* __restrict__ ctl,
                      const float* __restrict__ occ_box) {
    constexpr int kGroups = kCompactThreads / G;
    __shared__ float s_t[kGroups][kCompactCap];
    __shared__ uint32_t s_cnt[kGroups];
    __shared__ uint32_t s_excl;
    const Group<G> grp;
    const uint32_t n_alive = (uint32_t)ctl[kCtlAlive], n_step = (uint32_t)ctl[kCtlStep];
    if (ctl[kCtlRows] == 0) return;                      // finished: a queued no-op round
    const uint32_t active_blocks = div_up(n_alive, (uint32_t)kGroups);
    if (blockIdx.x >= active_blocks) return;
    const int gi = threadIdx.x / G, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x * kGroups + gi;
    const bool active = g < n_alive;
    Ray r = Ray{};
    float t0 = 0.f, far = 0.f;
    int index = 0;
    if (active) {
        index = __ldg(rays_alive + g);
        r = make_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
        far = clip_far_to_box(occ_box, r, fars[index]);
        t0 = rays_t[index];  // rounds >= 1 carry no noise (raymarching.cu:746 with noise = 0 leaves t_in unchanged)
    }
    bool inexact = false;
    const uint32_t cnt = march_group<G, true>(grp, p, r, grid, t0, far, n_step, active,
                                              [&](uint32_t rank, float s, float dt, const Probe&, float prev_after) {
                                                  s_t[gi][rank] = s;
                                                  const float t_after = f_add(s, dt), d1 = f_add(t_after, -prev_after);
                                                  inexact |= f_add(prev_after, d1) != t_after;
                                              });
    inexact = grp.ballot(inexact) != 0u;
    if (inexact) {  // rounds >= 1 only: see k_march_infer
        ctl[kCtlInexact] = 1;
        if (grp.gl == 0) {
            atomicAdd(ctl + 13, 1);
            uint8_t* flags = ctl_ptr<uint8_t>(ctl, kCtlFlagsPtr);
            if (flags) flags[index] = 1;
        }
    }
    if (grp.gl == 0) s_cnt[gi] = active ? cnt : 0u;
    __syncthreads();

    // ---- offsets: groups inside the block, then decoupled look-back over the blocks ----
    const uint32_t b = blockIdx.x;
    if (warp == 0) {
        uint32_t c = lane < kGroups ? s_cnt[lane] : 0u, incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane < kGroups) s_cnt[lane] = incl - c;  // exclusive prefix of the groups inside the block (their counts stay in registers: cnt)
        const uint32_t agg = __shfl_sync(kFull, incl, 31);
        uint32_t excl = 0;
        if (b > 0) {
            if (lane == 0) st_relaxed_u64(status + b, (1ull << 32) | agg);
            int idx = (int)b - 1;
            while (true) {
                const int j = idx - lane;
                unsigned long long st;
                do {
                    st = (j >= 0) ? ld_relaxed_u64(status + j) : (2ull << 32);
                } while (__any_sync(kFull, (st >> 32) == 0ull));
                const unsigned inc = __ballot_sync(kFull, (st >> 32) == 2ull);
                uint32_t x = (uint32_t)st;
                if (inc) {
                    const int first = __ffs(inc) - 1;
                    if (lane > first) x = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
                excl += x;
                if (inc) break;
                idx -= 32;
            }
        }
        if (lane == 0) {
            st_relaxed_u64(status + b, (2ull << 32) | (unsigned long long)(excl + agg));
            s_excl = excl;
        }
        if (b == active_blocks - 1) {  // the frame's rows of this round: published, and the rows up to the next multiple of 128 zeroed
            const uint32_t total = excl + agg, padded = (total + 127u) & ~127u;
            if (lane == 0) ctl[kCtlCompactRows] = (int)padded;
            for (uint32_t i = total * 3u + lane; i < padded * 3u; i += 32) { xyzs[i] = 0.f; dirs[i] = 0.f; }
            for (uint32_t i = total * 2u + lane; i < padded * 2u; i += 32) deltas[i] = 0.f;
            if (DISTILL)
                for (uint32_t i = total + lane; i < padded; i += 32) edit_occ[i] = 0;
        }
    }
    __syncthreads();
    if (!active) return;
    const uint32_t off = s_excl + s_cnt[gi];
    if (grp.gl == 0) { ray_off[g] = (int)off; ray_cnt[g] = (int)cnt; }

    // ---- the group writes its samples back to back: flat over the ray's floats, so that consecutive lanes store consecutive words
    // (positions recomputed from the parked t values, as k_march_train does) ----
    const float o3[3] = {r.ox, r.oy, r.oz};
    const float d3[3] = {r.dx, r.dy, r.dz};
    float* px = xyzs + (size_t)off * 3;
    float* pd = dirs + (size_t)off * 3;
    for (uint32_t e = grp.gl; e < cnt * 3u; e += G) {
        const uint32_t k = e / 3u, c = e - 3u * k;
        const float oc = c == 0 ? o3[0] : (c == 1 ? o3[1] : o3[2]);
        const float dc = c == 0 ? d3[0] : (c == 1 ? d3[1] : d3[2]);
        px[e] = f_clamp(f_fma(s_t[gi][k], dc, oc), p.neg_bound, p.bound);
        pd[e] = dc;
    }
    for (uint32_t k = grp.gl; k < cnt; k += G) {
        const float s = s_t[gi][k];
        const float dt = p.dt_const ? p.dt0 : march_dt(p, s);
        const float after = f_add(s, dt);
        float last = t0;
        if (k > 0) {
            const float sp = s_t[gi][k - 1];
            last = f_add(sp, p.dt_const ? p.dt0 : march_dt(p, sp));
        }
        const size_t row = (size_t)off + k;
        reinterpret_cast<float2*>(deltas)[row] = make_float2(dt, f_add(after, -last));
        if (DISTILL) {
            const Probe q = march_probe(p, r, s, dt);
            edit_occ[row] = (uint8_t)((__ldg(edit_grid + (q.index >> 3)) >> (q.index & 7u)) & 1u);
        }
    }
}

// composite_infer_ray (raymarching.cu:948-1035 / 1037-1142) over a ray's compact samples; also re-arms the marcher's status words.
// A ray's samples are consecutive rows, so a thread per ray reads with a stride of its neighbour's sample count: 32 sectors per load
// instruction (257 us for the 14 M samples of the lego frame's second round).  Here G lanes share a ray: they fetch G consecutive
// samples with one coalesced load each, and every lane then runs the reference's sequential recurrence over them, the operands
// handed round by shuffles -- the same operations in the same order on every lane, so the ray's state never leaves registers.
template <bool DISTILL, int G>
__global__ void __launch_bounds__(256)
k_composite_infer_compact(const float T_thresh, int* __restrict__ rays_alive, float* __restrict__ rays_t, const float* __restrict__ sigmas,
                          const float* __restrict__ rgbs, const float* __restrict__ deltas, float* __restrict__ weights_sum,
                          float* __restrict__ weights_edit_sum, float* __restrict__ depth, float* __restrict__ depth_edit,
                          const uint8_t* __restrict__ edit_occ, float* __restrict__ image, const int* __restrict__ ray_off,
                          const int* __restrict__ ray_cnt, unsigned long long* __restrict__ status, const uint32_t n_status,
                          const int* __restrict__ ctl) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n_status; i += nth) status[i] = 0ull;
    const uint32_t na = (uint32_t)ctl[kCtlAlive], n_step = (uint32_t)ctl[kCtlStep];
    int* ray_steps = ctl_ptr<int>(ctl, kCtlStepsPtr);
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? kFull : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));  // the rays of a warp loop independently
    for (uint32_t n = tid / G; n < na; n += nth / G) {
        const int index = rays_alive[n];
        const size_t o = (size_t)ray_off[n];
        const uint32_t count = (uint32_t)ray_cnt[n];
        const float* ps = sigmas + o;
        const float* pc = rgbs + o * 3;
        const float2* pl = reinterpret_cast<const float2*>(deltas) + o;
        const uint8_t* pe = DISTILL ? edit_occ + o : nullptr;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float weight_edit_sum = 0.f, d_edit = 0.f;
        if (DISTILL) { weight_edit_sum = weights_edit_sum[index]; d_edit = depth_edit[index]; }
        float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
        uint32_t step = 0;
        bool stopped = false;
        for (uint32_t base = 0; base < count && !stopped; base += G) {
            const uint32_t k = base + (uint32_t)gl;
            float m_sg = 0.f, m_cr = 0.f, m_cg = 0.f, m_cb = 0.f;
            float2 m_dl = make_float2(0.f, 0.f);
            int m_eo = 0;
            if (k < count) {
                m_sg = __ldg(ps + k);
                m_dl = __ldg(pl + k);
                m_cr = __ldg(pc + (size_t)k * 3); m_cg = __ldg(pc + (size_t)k * 3 + 1); m_cb = __ldg(pc + (size_t)k * 3 + 2);
                if (DISTILL) m_eo = pe[k];
            }
            const uint32_t m = count - base < (uint32_t)G ? count - base : (uint32_t)G;
            for (uint32_t j = 0; j < m; j++) {
                const float sg = __shfl_sync(gmask, m_sg, (int)j, G), dx = __shfl_sync(gmask, m_dl.x, (int)j, G),
                            dy = __shfl_sync(gmask, m_dl.y, (int)j, G);
                const float cr = __shfl_sync(gmask, m_cr, (int)j, G), cg = __shfl_sync(gmask, m_cg, (int)j, G),
                            cb = __shfl_sync(gmask, m_cb, (int)j, G);
                const float alpha = 1.0f - __expf(-sg * dx);
                const float T = f_add(1.0f, -weight_sum);
                const float weight = f_mul(alpha, T);
                weight_sum = f_add(weight_sum, weight);
                if (DISTILL) {
                    const int eo = __shfl_sync(gmask, m_eo, (int)j, G);
                    if (eo) {
                        weight_edit_sum = f_add(weight_edit_sum, weight);
                        d_edit = f_fma(weight, t, d_edit);
                    }
                }
                t = f_add(t, dy);
                d = f_fma(weight, t, d);
                r = f_fma(weight, cr, r);
                g = f_fma(weight, cg, g);
                b = f_fma(weight, cb, b);
                if (T < T_thresh) { stopped = true; break; }
                step++;
            }
        }
        if (gl == 0) {
            // the reference's slot layout ends a ray whose round came up short (dl.x == 0 in the next slot) or that hit T_thresh
            if (stopped || count < n_step) rays_alive[n] = -1; else rays_t[index] = t;
            if (ray_steps) ray_steps[index] += (int)step;
            if (DISTILL) { weights_edit_sum[index] = weight_edit_sum; depth_edit[index] = d_edit; }
            weights_sum[index] = weight_sum;
            depth[index] = d;
            image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
        }
    }
}

// =========================================================================================================
// a ray subset on a PRESCRIBED round schedule, all rounds in one pass (the fix-up of the "auto" render schedule)
// =========================================================================================================
// Where a ray's samples fall depends on the round boundaries only through rays_t, which composite_rays rebuilds at the end of
// every round as rays_t + sum of deltas[.][1] (raymarching.cu:1006) -- float additions of numbers the MARCHER produced.  The
// network's outputs only decide where the ray stops (T < T_thresh), never where its samples lie.  So for a given n_step sequence
// the whole ray can be marched in one go: round after round, each starting from the t the compositor would have rebuilt, until the
// ray leaves the volume or the sequence ends; the network then runs once over all those samples and the compositor walks them in
// order with the per-round arithmetic and stops at the reference's sample.  Same bits as running the rounds one by one
// (lnrf_render_rounds with nstep_seq) -- in 4 launches instead of 5 per round.
//   offsets == nullptr: count only (counts[ray] = samples up to the ray's exit); else write the ray's samples at offsets[ray].
//   caps (optional): most samples to march for each ray -- the caller knows roughly where a ray dies (the fast pass recorded it) and
//   spares the walk to the far end of the ray and the counting pass (offsets = prefix sum of the caps); a ray that has not died
//   within its cap must be redone without one.  occ_box: see clip_far_to_box.
template <bool DISTILL>
__global__ void __launch_bounds__(256)
k_march_prescribed(const uint32_t n_rays, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const float* __restrict__ nears, const float* __restrict__ fars, const MarchParams p,
                   const uint8_t* __restrict__ grid, const uint8_t* __restrict__ edit_grid, const int* __restrict__ seq,
                   const uint32_t seq_len, const int* __restrict__ offsets, int* __restrict__ counts, float* __restrict__ xyzs,
                   float* __restrict__ dirs, float* __restrict__ deltas, uint8_t* __restrict__ edit_occ, const int* __restrict__ caps,
                   const float* __restrict__ occ_box) {
    constexpr int G = kInferGroup;
    __shared__ float s_d1[256 / G][8];
    const Group<G> grp;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gi = threadIdx.x / G;
    bool active = g < n_rays;
    Ray r = Ray{};
    float t = 0.f, far = 0.f;
    size_t row = 0;
    if (active) {
        r = make_ray(rays_o + (size_t)g * 3, rays_d + (size_t)g * 3);
        far = clip_far_to_box(occ_box, r, fars[g]);
        t = nears[g];  // rays_t starts at near (renderer.py:349); no perturbation on this path
        if (offsets) row = (size_t)offsets[g];
    }
    const uint32_t cap = (active && caps) ? (uint32_t)max(caps[g], 0) : 0xffffffffu;
    uint32_t total = 0;
    for (uint32_t rd = 0; rd < seq_len; rd++) {
        if (!__any_sync(kFull, active)) break;
        const uint32_t n_step = (uint32_t)__ldg(seq + rd);  // <= 8
        if (active && total >= cap) active = false;
        const uint32_t want = n_step < cap - total ? n_step : cap - total;  // group-uniform (cap, total are the ray's)
        const uint32_t cnt = march_group<G, true>(
            grp, p, r, grid, t, far, active ? want : 0u, active, [&](uint32_t rank, float s, float dt, const Probe& q, float prev_after) {
                const float t_after = f_add(s, dt), d1 = f_add(t_after, -prev_after);
                s_d1[gi][rank] = d1;
                if (offsets) {
                    const size_t o = row + rank;
                    xyzs[o * 3] = q.x; xyzs[o * 3 + 1] = q.y; xyzs[o * 3 + 2] = q.z;
                    dirs[o * 3] = r.dx; dirs[o * 3 + 1] = r.dy; dirs[o * 3 + 2] = r.dz;
                    reinterpret_cast<float2*>(deltas)[o] = make_float2(dt, d1);
                    if (DISTILL) edit_occ[o] = (uint8_t)((__ldg(edit_grid + (q.index >> 3)) >> (q.index & 7u)) & 1u);
                }
            });
        __syncwarp();
        if (active) {
            float tc = t;  // what composite_rays leaves in rays_t for the next round: t += deltas[.][1], sample by sample
            for (uint32_t j = 0; j < cnt; j++) tc = f_add(tc, s_d1[gi][j]);
            t = tc;
            row += cnt;
            total += cnt;
            if (cnt < n_step) active = false;  // the ray ran out of samples inside this round: the compositor kills it
        }
        __syncwarp();
    }
    if (g < n_rays && grp.gl == 0) counts[g] = (int)total;
}

// raymarching.cu:948-1035 / 1037-1142 over all of a ray's rounds at once (thread per ray): the per-sample arithmetic of
// composite_infer_ray, t carried from sample to sample exactly as rays_t carries it from round to round.
template <bool DISTILL>
__global__ void __launch_bounds__(256)
k_composite_prescribed(const uint32_t n_rays, const float T_thresh, const int* __restrict__ offsets, const int* __restrict__ counts,
                       const float* __restrict__ nears, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                       const float* __restrict__ deltas, const uint8_t* __restrict__ edit_occ, float* __restrict__ weights_sum,
                       float* __restrict__ weights_edit_sum, float* __restrict__ depth, float* __restrict__ depth_edit,
                       float* __restrict__ image, int* __restrict__ ray_steps) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rays) return;
    const size_t o = (size_t)offsets[n];
    const uint32_t count = (uint32_t)counts[n];
    const float* ps = sigmas + o;
    const float* pc = rgbs + o * 3;
    const float2* pl = reinterpret_cast<const float2*>(deltas) + o;
    float t = nears[n];
    float weight_sum = 0.f, d = 0.f, weight_edit_sum = 0.f, d_edit = 0.f, r = 0.f, g = 0.f, b = 0.f;
    uint32_t step = 0;
    while (step < count) {
        const float2 dl = __ldg(pl + step);
        if (dl.x == 0.f) break;
        const float alpha = 1.0f - __expf(-__ldg(ps + step) * dl.x);
        const float T = f_add(1.0f, -weight_sum);
        const float weight = f_mul(alpha, T);
        weight_sum = f_add(weight_sum, weight);
        if (DISTILL && edit_occ[o + step]) {
            weight_edit_sum = f_add(weight_edit_sum, weight);
            d_edit = f_fma(weight, t, d_edit);
        }
        t = f_add(t, dl.y);
        d = f_fma(weight, t, d);
        r = f_fma(weight, __ldg(pc + step * 3), r);
        g = f_fma(weight, __ldg(pc + step * 3 + 1), g);
        b = f_fma(weight, __ldg(pc + step * 3 + 2), b);
        if (T < T_thresh) break;
        step++;
    }
    if (ray_steps) ray_steps[n] = (int)step;
    if (DISTILL) { weights_edit_sum[n] = weight_edit_sum; depth_edit[n] = d_edit; }
    weights_sum[n] = weight_sum;
    depth[n] = d;
    image[(size_t)n * 3] = r; image[(size_t)n * 3 + 1] = g; image[(size_t)n * 3 + 2] = b;
}

// start of a device-driven render: rays_alive = 0..n_rays-1, rays_t = nears, accumulators cleared, first round published
__global__ void __launch_bounds__(256)
k_render_begin(int* __restrict__ ctl, const uint32_t n_rays, const uint32_t max_steps, const uint32_t row_budget, const uint32_t step_cap,
               int* __restrict__ rays_alive,
               float* __restrict__ rays_t, const float* __restrict__ nears, float* __restrict__ weights_sum,
               float* __restrict__ depth, float* __restrict__ image, float* __restrict__ weights_edit_sum,
               float* __restrict__ depth_edit, int* __restrict__ ray_steps, uint8_t* __restrict__ ray_flags,
               const int* __restrict__ nstep_seq, const uint32_t nstep_len) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += gridDim.x * blockDim.x) {
        rays_alive[i] = (int)i;
        rays_t[i] = nears[i];
        if (ray_steps) ray_steps[i] = 0;
        if (ray_flags) ray_flags[i] = 0;
        weights_sum[i] = 0.f; depth[i] = 0.f;
        image[(size_t)i * 3] = 0.f; image[(size_t)i * 3 + 1] = 0.f; image[(size_t)i * 3 + 2] = 0.f;
        if (weights_edit_sum) { weights_edit_sum[i] = 0.f; depth_edit[i] = 0.f; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl[kCtlRays] = (int)n_rays;
        ctl[kCtlMaxSteps] = (int)max_steps;
        ctl[kCtlSteps] = 0;
        ctl[kCtlRounds] = 0;
        ctl[kCtlRounds + 1] = 0;
        ctl[kCtlRounds + 2] = 0;
        ctl[kCtlBudget] = (int)row_budget;
        ctl[kCtlStepCap] = (int)step_cap;
        ctl[kCtlInexact] = 0;
        ctl[13] = 0;
        ctl[14] = 0;
        ctl[kCtlSeqLen] = nstep_seq ? (int)nstep_len : 0;
        ctl_set_ptr(ctl, kCtlSeqPtr, nstep_seq);
        ctl_set_ptr(ctl, kCtlStepsPtr, ray_steps);
        ctl_set_ptr(ctl, kCtlFlagsPtr, ray_flags);
        ctl_set_round(ctl, n_rays);
    }
}


// ---- host launchers of the device-driven rounds (declared in render_core.cuh) ----
int render_begin_launch(int32_t* ctl, uint32_t n_rays, uint32_t max_steps, uint32_t row_budget, uint32_t step_cap, int32_t* rays_alive, float* rays_t,
                        const float* nears,
                        float* weights_sum, float* depth, float* image, float* weights_edit_sum, float* depth_edit, int32_t* ray_steps,
                        uint8_t* ray_flags, const int32_t* nstep_seq, uint32_t nstep_len, cudaStream_t st) {
    LNRF_REQUIRE(ctl && (n_rays == 0 || (rays_alive && rays_t && nears && weights_sum && depth && image)), "render_begin: null pointer");
    const uint32_t blocks = n_rays ? (div_up(n_rays, 256u) < (uint32_t)kNumSMs * 8u ? div_up(n_rays, 256u) : (uint32_t)kNumSMs * 8u) : 1u;
    LNRF_REQUIRE(row_budget >= n_rays, "render_begin: the row budget (%u) must cover one sample per ray (%u)", row_budget, n_rays);
    LNRF_REQUIRE(step_cap >= 1 && step_cap <= 64, "render_begin: samples per ray per round must be in [1, 64], got %u", step_cap);
    k_render_begin<<<blocks, 256, 0, st>>>(ctl, n_rays, max_steps, row_budget, step_cap, rays_alive, rays_t, nears, weights_sum, depth, image, weights_edit_sum,
                                           depth_edit, ray_steps, ray_flags, nstep_seq, nstep_len);
    LNRF_LAUNCH_CHECK("render_begin");
    return LNRF_OK;
}

int march_infer_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, const float* rays_t,
                           const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                           uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* fars, float* xyzs, float* dirs,
                           float* deltas, uint8_t* edit_occ, const float* noises, bool first, const float* occ_box, cudaStream_t st) {
    const char* who = distill ? "render_rounds(march_distill)" : "render_rounds(march)";
    if (int e = check_march_common(C, H, max_steps, who)) return e;
    if (n_rays_cap == 0) return LNRF_OK;
    LNRF_REQUIRE(ctl && rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas, "%s: null pointer", who);
    const MarchParams p = march_params_env(bound, dt_gamma, max_steps, C, H);
    // the grid covers the ray capacity (a round has at most n_rays_cap + 128 groups); blocks past the round's groups read the
    // control block and leave (cheaper than a grid-stride loop over a persistent grid: 97 vs 115 us per steady-state round)
    const uint32_t cap_blocks = 0xffffffffu;
#define LNRF_MARCH_DEV_(DD, GG, LO, HI)                                                                                               \
    {                                                                                                                                 \
        const uint32_t want = div_up(n_rays_cap + 128u, 256u / GG);                                                                   \
        k_march_infer<DD, GG><<<want < cap_blocks ? want : cap_blocks, 256, 0, st>>>(0u, 1u, rays_alive, rays_t, rays_o, rays_d, p, grid, \
                                                                                      edit_grid, fars, xyzs, dirs, deltas, edit_occ,   \
                                                                                      noises, 0u, 0u, ctl, LO, HI, occ_box);           \
        LNRF_LAUNCH_CHECK(who);                                                                                                       \
    }
    // Group width, measured per round on the 640 000-ray lego frame (profiles/r1g_render.txt): first round (every ray walks
    // from `near` through ~440 empty sequence members, one voxel ~ 4.6 members) 4 lanes 2.08 ms, 8 lanes 2.22 ms, 32 lanes
    // 3.69 ms -- a window wider than a voxel only adds resolve iterations and idle probes; later rounds (rays continue inside
    // or next to occupied cells, 2-3 samples wanted) 8 lanes 97 us, 4 lanes 150-200 us -- a second window costs more than
    // four idle lanes.
    if (first) {
        static const int first_g = [] { const char* e = getenv("LNRF_MARCH_FIRST_G"); return e ? atoi(e) : 4; }();  // A/B switch
        if (first_g == 1) { if (distill) LNRF_MARCH_DEV_(true, 1, 1, 64) else LNRF_MARCH_DEV_(false, 1, 1, 64) }
        else if (first_g == 2) { if (distill) LNRF_MARCH_DEV_(true, 2, 1, 64) else LNRF_MARCH_DEV_(false, 2, 1, 64) }
        else { if (distill) LNRF_MARCH_DEV_(true, 4, 1, 64) else LNRF_MARCH_DEV_(false, 4, 1, 64) }
    } else {
        if (distill) LNRF_MARCH_DEV_(true, 8, 1, 64) else LNRF_MARCH_DEV_(false, 8, 1, 64)
    }
#undef LNRF_MARCH_DEV_
    return LNRF_OK;
}

int composite_infer_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                               const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* weights_edit_sum,
                               float* depth, float* depth_edit, const uint8_t* edit_occ, float* image, cudaStream_t st) {
    if (n_rays_cap == 0) return LNRF_OK;
    const uint32_t want = div_up(n_rays_cap, 256u), cap_blocks = (uint32_t)kNumSMs * 8u;
    const uint32_t blocks = want < cap_blocks ? want : cap_blocks;
    if (distill)
        k_composite_infer<true><<<blocks, 256, 0, st>>>(0u, 1u, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum,
                                                        weights_edit_sum, depth, depth_edit, edit_occ, image, ctl);
    else
        k_composite_infer<false><<<blocks, 256, 0, st>>>(0u, 1u, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, nullptr,
                                                         depth, nullptr, nullptr, image, ctl);
    LNRF_LAUNCH_CHECK("render_rounds(composite)");
    return LNRF_OK;
}

// compact rounds: scratch_m = (n_rays_cap / 4 + 2) status words (one per marcher block at most), then ray_off[n_rays_cap], ray_cnt[n_rays_cap]
size_t march_compact_scratch_bytes(uint32_t n_rays_cap) {
    return sizeof(unsigned long long) * ((size_t)div_up(n_rays_cap, 4u) + 2) + 2 * sizeof(int) * (size_t)n_rays_cap;
}
static void march_compact_carve(void* scratch_m, uint32_t n_rays_cap, unsigned long long** status, uint32_t* n_status, int** off, int** cnt) {
    *status = reinterpret_cast<unsigned long long*>(scratch_m);
    *n_status = div_up(n_rays_cap, 4u) + 2u;   // one word per marcher block; 4 rays per block with 32 lanes per ray
    *off = reinterpret_cast<int*>(*status + *n_status);
    *cnt = *off + n_rays_cap;
}

int march_infer_compact_dev_launch(bool distill, int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, const float* rays_t,
                                   const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                   uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* fars, float* xyzs, float* dirs,
                                   float* deltas, uint8_t* edit_occ, void* scratch_m, const float* occ_box, cudaStream_t st) {
    const char* who = distill ? "render_rounds(march_distill, compact)" : "render_rounds(march, compact)";
    if (int e = check_march_common(C, H, max_steps, who)) return e;
    if (n_rays_cap == 0) return LNRF_OK;
    LNRF_REQUIRE(ctl && rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && scratch_m, "%s: null pointer", who);
    const MarchParams p = march_params_env(bound, dt_gamma, max_steps, C, H);
    unsigned long long* status; uint32_t n_status; int *off, *cnt;
    march_compact_carve(scratch_m, n_rays_cap, &status, &n_status, &off, &cnt);
    // lanes per ray (LNRF_COMPACT_G = 4 | 8 | 16 | 32, A/B switch; measured lego / bonsai frame: 8 lanes 11.2 / 28.1 ms, 16: 11.7 / 31.9, 32: 12.5 / 39.4): a window of G sequence members per iteration
    // Default: 8 lanes, 4 when the scene has three or more cascades -- there most of a ray's walk crosses cells much longer than a
    // window of sequence members (cell 0.25 against 0.027 for 8 members on cascade 4 of the bonsai shape), so narrow groups idle less:
    // bonsai frame 28.3 -> 25.4 ms with 4 lanes, lego (one cascade) 11.4 -> 11.5.
    static const int forced = [] { const char* e = getenv("LNRF_COMPACT_G"); const int v = e ? atoi(e) : 0; return (v == 4 || v == 8 || v == 16 || v == 32) ? v : 0; }();
    const int cg = forced ? forced : (C >= 3 ? 4 : kCompactG);
#define LNRF_MIC_(DD, GG)                                                                                                                  \
    k_march_infer_compact<DD, GG><<<div_up(n_rays_cap, (uint32_t)kCompactThreads / GG), kCompactThreads, 0, st>>>(                          \
        rays_alive, rays_t, rays_o, rays_d, p, grid, DD ? edit_grid : nullptr, fars, xyzs, dirs, deltas, DD ? edit_occ : nullptr, off, cnt, status, ctl, occ_box)
    if (distill) { if (cg == 32) LNRF_MIC_(true, 32); else if (cg == 16) LNRF_MIC_(true, 16); else if (cg == 4) LNRF_MIC_(true, 4); else LNRF_MIC_(true, 8); }
    else { if (cg == 32) LNRF_MIC_(false, 32); else if (cg == 16) LNRF_MIC_(false, 16); else if (cg == 4) LNRF_MIC_(false, 4); else LNRF_MIC_(false, 8); }
#undef LNRF_MIC_
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

int composite_infer_compact_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                                       const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                       float* weights_edit_sum, float* depth, float* depth_edit, const uint8_t* edit_occ, float* image,
                                       void* scratch_m, cudaStream_t st) {
    if (n_rays_cap == 0) return LNRF_OK;
    unsigned long long* status; uint32_t n_status; int *off, *cnt;
    march_compact_carve(scratch_m, n_rays_cap, &status, &n_status, &off, &cnt);
    static const int cg = [] { const char* e = getenv("LNRF_COMPOSITE_INFER_G"); const int v = e ? atoi(e) : 4; return (v == 1 || v == 4 || v == 8 || v == 16) ? v : 4; }();  // measured on the lego frame: 1 lane 11.68 ms, 4: 11.22, 8: 11.30, 16: 11.68
#define LNRF_CIC_(DD, GG)                                                                                                               \
    {                                                                                                                                   \
        const uint32_t want = div_up(n_rays_cap, 256u / GG), cap_blocks = (uint32_t)kNumSMs * 8u;                                       \
        k_composite_infer_compact<DD, GG><<<want < cap_blocks ? want : cap_blocks, 256, 0, st>>>(                                       \
            T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, DD ? weights_edit_sum : nullptr, depth, DD ? depth_edit : nullptr, \
            DD ? edit_occ : nullptr, image, off, cnt, status, n_status, ctl);                                                           \
    }
    if (distill) {
        if (cg == 1) LNRF_CIC_(true, 1) else if (cg == 16) LNRF_CIC_(true, 16) else if (cg == 8) LNRF_CIC_(true, 8) else LNRF_CIC_(true, 4)
    } else {
        if (cg == 1) LNRF_CIC_(false, 1) else if (cg == 16) LNRF_CIC_(false, 16) else if (cg == 8) LNRF_CIC_(false, 8) else LNRF_CIC_(false, 4)
    }
#undef LNRF_CIC_
    LNRF_LAUNCH_CHECK("render_rounds(composite, compact)");
    return LNRF_OK;
}

int compact_alive_dev_launch(int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, int32_t* out, void* scratch,
                             size_t scratch_bytes, cudaStream_t st) {
    if (n_rays_cap == 0) return LNRF_OK;
    if (!scratch || scratch_bytes < lnrf_compact_alive_scratch_bytes(n_rays_cap)) {
        set_error("render_rounds: compaction scratch too small (%zu < %zu)", scratch_bytes, lnrf_compact_alive_scratch_bytes(n_rays_cap));
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    k_compact_alive<<<div_up(n_rays_cap, 1024u), 1024, 0, st>>>(rays_alive, n_rays_cap, out, nullptr,
                                                                reinterpret_cast<unsigned long long*>(scratch), ctl);
    LNRF_LAUNCH_CHECK("render_rounds(compact)");
    return LNRF_OK;
}

}  // namespace lnrf

// =========================================================================================================
// C ABI
// =========================================================================================================
using namespace lnrf;

static inline cudaStream_t S(lnrf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int lnrf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                            float* nears, float* fars, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(rays_o && rays_d && aabb && nears && fars, "near_far_from_aabb: null pointer");
    k_near_far<<<div_up(N, 256u), 256, 0, S(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    LNRF_LAUNCH_CHECK("near_far_from_aabb");
    return LNRF_OK;
}

int lnrf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(rays_o && rays_d && coords, "sph_from_ray: null pointer");
    k_sph_from_ray<<<div_up(N, 256u), 256, 0, S(stream)>>>(rays_o, rays_d, radius, N, coords);
    LNRF_LAUNCH_CHECK("sph_from_ray");
    return LNRF_OK;
}

int lnrf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(coords && indices, "morton3D: null pointer");
    k_morton3d<<<div_up(N, 256u), 256, 0, S(stream)>>>(coords, N, indices);
    LNRF_LAUNCH_CHECK("morton3D");
    return LNRF_OK;
}

int lnrf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(coords && indices, "morton3D_invert: null pointer");
    k_morton3d_invert<<<div_up(N, 256u), 256, 0, S(stream)>>>(indices, N, coords);
    LNRF_LAUNCH_CHECK("morton3D_invert");
    return LNRF_OK;
}

int lnrf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(grid && bitfield, "packbits: null pointer");
    const uint32_t words = div_up(N, 4u);
    k_packbits<<<div_up(words, 256u), 256, 0, S(stream)>>>(grid, N, density_thresh, bitfield);
    LNRF_LAUNCH_CHECK("packbits");
    return LNRF_OK;
}

size_t lnrf_march_rays_train_scratch_bytes(uint32_t N) { return sizeof(unsigned long long) * ((size_t)N + 2); }

static int check_march_common(uint32_t C, uint32_t H, uint32_t max_steps, const char* who) {
    LNRF_REQUIRE(C >= 1 && C <= 8, "%s: cascade count C=%u out of range [1,8]", who, C);
    LNRF_REQUIRE(H >= 2 && H <= 1024, "%s: grid size H=%u out of range", who, H);
    // the reference evaluates the bit index in float (H3 is a float, raymarching.cu:339,378): exact only below 2^24
    LNRF_REQUIRE((uint64_t)C * H * H * H <= (1ull << 24), "%s: C*H^3 = %llu exceeds 2^24 (index is computed in float)", who,
                 (unsigned long long)C * H * H * H);
    LNRF_REQUIRE(max_steps >= 1, "%s: max_steps must be >= 1", who);
    return LNRF_OK;
}

int lnrf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                          const float* noises, void* scratch, size_t scratch_bytes, lnrf_stream_t stream) {
    return lnrf_march_rays_train_clipped(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                                         counter, noises, nullptr, scratch, scratch_bytes, stream);
}

int lnrf_march_rays_train_clipped(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                                  uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                                  const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                                  const float* noises, const float* occupied_box, void* scratch, size_t scratch_bytes,
                                  lnrf_stream_t stream) {
    if (int e = check_march_common(C, H, max_steps, "march_rays_train")) return e;
    LNRF_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises, "march_rays_train: null pointer");
    LNRF_REQUIRE(M == 0 || (xyzs && dirs && deltas), "march_rays_train: null sample buffer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(xyzs) & 15) == 0 && (reinterpret_cast<uintptr_t>(dirs) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(deltas) & 15) == 0,
                 "march_rays_train: sample buffers must be 16-byte aligned");
    if (N == 0) return LNRF_OK;
    if (scratch_bytes < lnrf_march_rays_train_scratch_bytes(N) || !scratch) {
        set_error("march_rays_train: scratch too small (%zu < %zu)", scratch_bytes, lnrf_march_rays_train_scratch_bytes(N));
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    const MarchParams p = march_params_env(bound, dt_gamma, max_steps, C, H);
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(scratch);
    uint32_t nblocks;
    if ((size_t)max_steps * 4 * sizeof(float) <= 96 * 1024) {
        constexpr int W = 4;
        const size_t smem = (size_t)W * max_steps * sizeof(float);
        nblocks = div_up(N, (uint32_t)W);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_march_train<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(e, "march_rays_train: smem attribute");
        }
        k_march_train<W><<<nblocks, W * 32, smem, S(stream)>>>(rays_o, rays_d, grid, p, N, M, nears, fars, noises, xyzs, dirs,
                                                               deltas, rays, counter, sc, occupied_box);
    } else {
        constexpr int W = 1;
        const size_t smem = (size_t)W * max_steps * sizeof(float);
        LNRF_REQUIRE(smem <= 227 * 1024, "march_rays_train: max_steps=%u needs %zu B of shared memory (> 227 KiB)", max_steps, smem);
        nblocks = N;
        cudaError_t e = cudaFuncSetAttribute(k_march_train<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "march_rays_train: smem attribute");
        k_march_train<W><<<nblocks, W * 32, smem, S(stream)>>>(rays_o, rays_d, grid, p, N, M, nears, fars, noises, xyzs, dirs,
                                                               deltas, rays, counter, sc, occupied_box);
    }
    LNRF_LAUNCH_CHECK("march_rays_train");
    k_march_train_tail<<<kNumSMs, 256, 0, S(stream)>>>(sc, nblocks, xyzs, dirs, deltas, M);
    LNRF_LAUNCH_CHECK("march_rays_train(tail)");
    return LNRF_OK;
}

// blocks of the training compositors: one warp per ray up to 8 blocks per SM, grid-stride beyond
static uint32_t composite_loss_blocks(uint32_t N) {
    const uint32_t want = div_up(N, 8u), cap = (uint32_t)kNumSMs * 8u;
    return want < cap ? want : cap;
}

int lnrf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                      uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth, float* image,
                                      lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(rays && weights_sum && depth && image, "composite_rays_train_forward: null pointer");
    LNRF_REQUIRE(M == 0 || (sigmas && rgbs && deltas), "composite_rays_train_forward: null sample buffer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_rays_train_forward: deltas must be 8-byte aligned");
    // one block per 8 rays: a capped grid-stride grid measured SLOWER here at 65 536 rays (46 vs 31-34 us)
    k_composite_train_fwd<<<div_up(N, 8u), 256, 0, S(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    LNRF_LAUNCH_CHECK("composite_rays_train_forward");
    return LNRF_OK;
}

int lnrf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                       const float* rgbs, const float* deltas, const int32_t* rays, const float* weights_sum,
                                       const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                       float* grad_rgbs, int zero_fill, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(grad_weights_sum && grad_image && rays && weights_sum && image, "composite_rays_train_backward: null pointer");
    LNRF_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs), "composite_rays_train_backward: null sample buffer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_rays_train_backward: deltas must be 8-byte aligned");
    if (zero_fill)
        k_composite_train_bwd<true, false><<<div_up(N, 8u), 256, 0, S(stream)>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays,
                                                                                 weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs,
                                                                                 LossArgs{});
    else
        k_composite_train_bwd<false, false><<<div_up(N, 8u), 256, 0, S(stream)>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays,
                                                                                  weights_sum, image, M, N, T_thresh, grad_sigmas,
                                                                                  grad_rgbs, LossArgs{});
    LNRF_LAUNCH_CHECK("composite_rays_train_backward");
    return LNRF_OK;
}

size_t lnrf_composite_loss_scratch_bytes(uint32_t N) { return ((size_t)div_up(N, 8u) + 4u) * sizeof(float); }

int lnrf_composite_loss_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                      const float* gt_rgb, const float* bg_rgb, float bg_scalar, const float* nears,
                                      const float* fars, uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                      float* image, float* image_raw, float* loss, void* scratch, size_t scratch_bytes,
                                      lnrf_stream_t stream) {
    LNRF_REQUIRE(N > 0, "composite_loss_train_forward: the mean over zero rays is undefined");
    LNRF_REQUIRE(rays && gt_rgb && weights_sum && depth && image && image_raw && loss && scratch, "composite_loss_train_forward: null pointer");
    LNRF_REQUIRE((nears == nullptr) == (fars == nullptr), "composite_loss_train_forward: nears and fars go together");
    LNRF_REQUIRE(M == 0 || (sigmas && rgbs && deltas), "composite_loss_train_forward: null sample buffer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_loss_train_forward: deltas must be 8-byte aligned");
    LNRF_REQUIRE(scratch_bytes >= lnrf_composite_loss_scratch_bytes(N), "composite_loss_train_forward: scratch too small");
    LossArgs la{};
    la.gt = gt_rgb; la.bg = bg_rgb; la.bg_scalar = bg_scalar; la.nears = nears; la.fars = fars; la.image_raw = image_raw;
    la.ticket = reinterpret_cast<unsigned int*>(scratch);
    la.partial = reinterpret_cast<float*>(scratch) + 4;
    la.loss = loss;
    launch_pdl(k_composite_loss_fwd, composite_loss_blocks(N), 256, 0, S(stream), sigmas, rgbs, deltas, rays, M, N, T_thresh, la, weights_sum, depth, image);
    LNRF_LAUNCH_CHECK("composite_loss_train_forward");
    return LNRF_OK;
}

int lnrf_composite_loss_train_forward_backward(const float* grad_loss, const float* sigmas, const float* rgbs, const float* deltas,
                                               const int32_t* rays, const float* gt_rgb, const float* bg_rgb, float bg_scalar,
                                               const float* nears, const float* fars, uint32_t M, uint32_t N, float T_thresh,
                                               float* weights_sum, float* depth, float* image, float* image_raw, float* loss,
                                               float* grad_sigmas, float* grad_rgbs, void* scratch, size_t scratch_bytes,
                                               lnrf_stream_t stream) {
    const char* who = "composite_loss_train_forward_backward";
    LNRF_REQUIRE(N > 0, "%s: the mean over zero rays is undefined", who);
    LNRF_REQUIRE(grad_loss && rays && gt_rgb && weights_sum && depth && image && image_raw && loss && scratch, "%s: null pointer", who);
    LNRF_REQUIRE((nears == nullptr) == (fars == nullptr), "%s: nears and fars go together", who);
    LNRF_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs), "%s: null sample buffer", who);
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "%s: deltas must be 8-byte aligned", who);
    LNRF_REQUIRE(scratch_bytes >= lnrf_composite_loss_scratch_bytes(N), "%s: scratch too small", who);
    LossArgs la{};
    la.gt = gt_rgb; la.bg = bg_rgb; la.bg_scalar = bg_scalar; la.nears = nears; la.fars = fars; la.image_raw = image_raw;
    la.ticket = reinterpret_cast<unsigned int*>(scratch);
    la.partial = reinterpret_cast<float*>(scratch) + 4;
    la.loss = loss;
    la.grad_loss = grad_loss;
    // One launch while the batch is about one wave of blocks (the 4096-ray training step: latency-bound, the second pass over the
    // ray's samples hits L1).  Large batches are throughput-bound and the two-pass warps hold their registers twice as long:
    // measured at 65 536 rays 191 us fused against 62 + 68 us for the two kernels -- which produce the same bits.
    if (N <= 8192u) {
        launch_pdl(k_composite_loss_fwd_bwd, composite_loss_blocks(N), 256, 0, S(stream), sigmas, rgbs, deltas, rays, M, N, T_thresh, la, weights_sum, depth,
                   image, grad_sigmas, grad_rgbs);
    } else {
        k_composite_loss_fwd<<<composite_loss_blocks(N), 256, 0, S(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh, la, weights_sum, depth, image);
        LNRF_LAUNCH_CHECK(who);
        k_composite_train_bwd<true, true><<<div_up(N, 8u), 256, 0, S(stream)>>>(nullptr, nullptr, sigmas, rgbs, deltas, rays, weights_sum,
                                                                                image, M, N, T_thresh, grad_sigmas, grad_rgbs, la);
    }
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

int lnrf_composite_loss_train_backward(const float* grad_loss, const float* sigmas, const float* rgbs, const float* deltas,
                                       const int32_t* rays, const float* gt_rgb, const float* bg_rgb, float bg_scalar,
                                       const float* weights_sum, const float* image, const float* image_raw, uint32_t M, uint32_t N,
                                       float T_thresh, float* grad_sigmas, float* grad_rgbs, lnrf_stream_t stream) {
    if (N == 0) return LNRF_OK;
    LNRF_REQUIRE(grad_loss && rays && gt_rgb && weights_sum && image && image_raw, "composite_loss_train_backward: null pointer");
    LNRF_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs), "composite_loss_train_backward: null sample buffer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_loss_train_backward: deltas must be 8-byte aligned");
    LossArgs la{};
    la.gt = gt_rgb; la.bg = bg_rgb; la.bg_scalar = bg_scalar; la.image_raw = const_cast<float*>(image_raw); la.grad_loss = grad_loss;
    launch_pdl(k_composite_train_bwd<true, true>, div_up(N, 8u), 256, 0, S(stream), nullptr, nullptr, sigmas, rgbs, deltas, rays, weights_sum,
               image, M, N, T_thresh, grad_sigmas, grad_rgbs, la);
    LNRF_LAUNCH_CHECK("composite_loss_train_backward");
    return LNRF_OK;
}

static int march_infer_launch(bool distill, uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                              const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                              uint32_t C, uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* fars,
                              float* xyzs, float* dirs, float* deltas, uint8_t* edit_occ, const float* noises,
                              uint32_t M_rows, cudaStream_t st) {
    const char* who = distill ? "march_rays_distill" : "march_rays";
    if (int e = check_march_common(C, H, max_steps, who)) return e;
    LNRF_REQUIRE(n_step >= 1, "%s: n_step must be >= 1", who);
    LNRF_REQUIRE((uint64_t)n_alive * n_step <= M_rows, "%s: M_rows=%u smaller than n_alive*n_step", who, M_rows);
    if (M_rows == 0) return LNRF_OK;
    LNRF_REQUIRE(xyzs && dirs && deltas && (!distill || (edit_occ && edit_grid)), "%s: null output", who);
    LNRF_REQUIRE(n_alive == 0 || (rays_alive && rays_t && rays_o && rays_d && grid && fars && noises), "%s: null input", who);
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "%s: deltas must be 8-byte aligned", who);
    const MarchParams p = march_params_env(bound, dt_gamma, max_steps, C, H);
    const uint32_t n_groups = div_up(M_rows, n_step);
#define LNRF_MARCH_INFER_(DD, GG)                                                                                                        \
    k_march_infer<DD, GG><<<div_up(n_groups, 256u / GG), 256, 0, st>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, p, grid,   \
                                                                       edit_grid, fars, xyzs, dirs, deltas, edit_occ, noises, M_rows,  \
                                                                       n_groups, nullptr, 0, 0)
    if (distill) LNRF_MARCH_INFER_(true, 8); else LNRF_MARCH_INFER_(false, 8);
#undef LNRF_MARCH_INFER_
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

int lnrf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                    const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                    const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                    const float* noises, uint32_t M_rows, lnrf_stream_t stream) {
    (void)nears;  // loaded but unused by the reference kernel too (raymarching.cu:737)
    return march_infer_launch(false, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                              nullptr, fars, xyzs, dirs, deltas, nullptr, noises, M_rows, S(stream));
}

int lnrf_march_rays_distill(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                            const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t C, uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* nears,
                            const float* fars, float* xyzs, float* dirs, float* deltas, uint8_t* edit_occ,
                            const float* noises, uint32_t M_rows, lnrf_stream_t stream) {
    (void)nears;
    return march_infer_launch(true, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                              edit_grid, fars, xyzs, dirs, deltas, edit_occ, noises, M_rows, S(stream));
}

int lnrf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                        const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                        float* image, lnrf_stream_t stream) {
    if (n_alive == 0) return LNRF_OK;
    LNRF_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, "composite_rays: null pointer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_rays: deltas must be 8-byte aligned");
    k_composite_infer<false><<<div_up(n_alive, 256u), 256, 0, S(stream)>>>(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs,
                                                                           deltas, weights_sum, nullptr, depth, nullptr, nullptr, image,
                                                                           nullptr);
    LNRF_LAUNCH_CHECK("composite_rays");
    return LNRF_OK;
}

int lnrf_composite_rays_distill(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                float* weights_edit_sum, float* depth, float* depth_edit, const uint8_t* edit_occ, float* image,
                                lnrf_stream_t stream) {
    if (n_alive == 0) return LNRF_OK;
    LNRF_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && weights_edit_sum && depth && depth_edit &&
                     edit_occ && image,
                 "composite_rays_distill: null pointer");
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "composite_rays_distill: deltas must be 8-byte aligned");
    k_composite_infer<true><<<div_up(n_alive, 256u), 256, 0, S(stream)>>>(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs,
                                                                          deltas, weights_sum, weights_edit_sum, depth, depth_edit,
                                                                          edit_occ, image, nullptr);
    LNRF_LAUNCH_CHECK("composite_rays_distill");
    return LNRF_OK;
}

int lnrf_march_rays_prescribed(uint32_t n_rays, const float* rays_o, const float* rays_d, const float* nears, const float* fars, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const uint8_t* edit_grid,
                               const int32_t* nstep_seq, uint32_t nstep_len, const int32_t* offsets, int32_t* counts, float* xyzs,
                               float* dirs, float* deltas, uint8_t* edit_occ, const int32_t* caps, const float* occupied_box,
                               lnrf_stream_t stream) {
    const char* who = "march_rays_prescribed";
    if (int e = check_march_common(C, H, max_steps, who)) return e;
    if (n_rays == 0) return LNRF_OK;
    LNRF_REQUIRE(rays_o && rays_d && nears && fars && grid && nstep_seq && counts, "%s: null pointer", who);
    LNRF_REQUIRE(!offsets || (xyzs && dirs && deltas && (!edit_grid || edit_occ)), "%s: writing pass without sample buffers", who);
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "%s: deltas must be 8-byte aligned", who);
    const MarchParams p = march_params_env(bound, dt_gamma, max_steps, C, H);
    const uint32_t blocks = div_up(n_rays, 256u / (uint32_t)kInferGroup);
    if (edit_grid)
        k_march_prescribed<true><<<blocks, 256, 0, S(stream)>>>(n_rays, rays_o, rays_d, nears, fars, p, grid, edit_grid, nstep_seq, nstep_len, offsets,
                                                               counts, xyzs, dirs, deltas, edit_occ, caps, occupied_box);
    else
        k_march_prescribed<false><<<blocks, 256, 0, S(stream)>>>(n_rays, rays_o, rays_d, nears, fars, p, grid, nullptr, nstep_seq, nstep_len, offsets,
                                                                counts, xyzs, dirs, deltas, nullptr, caps, occupied_box);
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

int lnrf_composite_rays_prescribed(uint32_t n_rays, float T_thresh, const int32_t* offsets, const int32_t* counts, const float* nears,
                                   const float* sigmas, const float* rgbs, const float* deltas, const uint8_t* edit_occ, float* weights_sum,
                                   float* weights_edit_sum, float* depth, float* depth_edit, float* image, int32_t* ray_steps,
                                   lnrf_stream_t stream) {
    const char* who = "composite_rays_prescribed";
    if (n_rays == 0) return LNRF_OK;
    LNRF_REQUIRE(offsets && counts && nears && sigmas && rgbs && deltas && weights_sum && depth && image, "%s: null pointer", who);
    LNRF_REQUIRE(!edit_occ || (weights_edit_sum && depth_edit), "%s: distillation needs weights_edit_sum / depth_edit", who);
    LNRF_REQUIRE((reinterpret_cast<uintptr_t>(deltas) & 7) == 0, "%s: deltas must be 8-byte aligned", who);
    if (edit_occ)
        k_composite_prescribed<true><<<div_up(n_rays, 256u), 256, 0, S(stream)>>>(n_rays, T_thresh, offsets, counts, nears, sigmas, rgbs, deltas, edit_occ,
                                                                                  weights_sum, weights_edit_sum, depth, depth_edit, image, ray_steps);
    else
        k_composite_prescribed<false><<<div_up(n_rays, 256u), 256, 0, S(stream)>>>(n_rays, T_thresh, offsets, counts, nears, sigmas, rgbs, deltas, nullptr,
                                                                                   weights_sum, nullptr, depth, nullptr, image, ray_steps);
    LNRF_LAUNCH_CHECK(who);
    return LNRF_OK;
}

size_t lnrf_compact_alive_scratch_bytes(uint32_t n_alive) { return sizeof(unsigned long long) * ((size_t)div_up(n_alive, 1024u) + 1); }

int lnrf_compact_alive(const int32_t* rays_alive, uint32_t n_alive, int32_t* out, int32_t* n_out, void* scratch,
                       size_t scratch_bytes, lnrf_stream_t stream) {
    LNRF_REQUIRE(n_out, "compact_alive: null counter");
    if (n_alive == 0) {
        cudaError_t e = cudaMemsetAsync(n_out, 0, sizeof(int32_t), S(stream));
        if (e != cudaSuccess) return cuda_fail(e, "compact_alive");
        return LNRF_OK;
    }
    LNRF_REQUIRE(rays_alive && out, "compact_alive: null pointer");
    if (!scratch || scratch_bytes < lnrf_compact_alive_scratch_bytes(n_alive)) {
        set_error("compact_alive: scratch too small (%zu < %zu)", scratch_bytes, lnrf_compact_alive_scratch_bytes(n_alive));
        return LNRF_ERR_SCRATCH_TOO_SMALL;
    }
    k_compact_alive<<<div_up(n_alive, 1024u), 1024, 0, S(stream)>>>(rays_alive, n_alive, out, n_out,
                                                                   reinterpret_cast<unsigned long long*>(scratch), nullptr);
    LNRF_LAUNCH_CHECK("compact_alive");
    return LNRF_OK;
}

}  // extern "C"
