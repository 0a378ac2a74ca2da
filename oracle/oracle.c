/*
 * oracle.c -- CPU restatement of LAENeRF's ray-marched NeRF step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA kernels in laenerf_b200/csrc.  It is never linked into,
 * imported by, or called from the product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  Every function cites the reference lines it restates
 * (paths relative to /root/reference).
 *
 * Pinning: the reference ships no golden vectors, so this oracle is pinned against the
 * outputs of the reference's own CUDA extensions compiled for sm_100a (oracle/build_ref.py) and run on a
 * B200 (oracle/gen_golden.py -> tests/golden/ref_*.npz).  tests/test_oracle_golden.py checks every function
 * below against those vectors.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).  Contraction is OFF so
 * that every fused multiply-add is an explicit fmaf() placed exactly where nvcc contracts the reference
 * source (verified from the reference SASS).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* helpers: raymarching/src/raymarching.cu:19-81                                                      */
/* ------------------------------------------------------------------------------------------------ */

static const float TWO_SQRT3 = 3.4641015529632568f; /* 2 * SQRT3() evaluated in float, :19,345 */

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); } /* :34-36 */
static inline float signf1(float x) { return copysignf(1.0f, x); }                          /* :30-32 */

/* :42-47.  frexpf(0) gives exponent 0 on both CUDA and glibc. */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)e));
}

/* :49-54.  dt*H is a float product; the *0.5 (double literal) is exact. */
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (dt * H) * 0.5f;
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)e));
}

static inline uint32_t expand_bits(uint32_t v) { /* :56-63 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D_1(uint32_t x, uint32_t y, uint32_t z) { /* :65-71 */
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_1(uint32_t x) { /* :73-81 */
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* ------------------------------------------------------------------------------------------------ */
/* utilities: raymarching.cu:91-300                                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* kernel_near_far_from_aabb, :91-145 */
ORC_API void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                    float min_near, float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1.0f / dx, rdy = 1.0f / dy, rdz = 1.0f / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
        if (near > far) { float c = near; near = far; far = c; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* kernel_sph_from_ray, :162-198 (tolerance-level: atan2f/sqrtf on device are not bit-identical to libm) */
ORC_API void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
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
}

/* kernel_morton3D :214-226, kernel_morton3D_invert :237-254 */
ORC_API void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3D_1((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
ORC_API void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n]; /* signed >> as in the reference (:249-253) */
        coords[n * 3] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3D_invert_1((uint32_t)(ind >> 2));
    }
}

/* kernel_packbits, :267-289.  N = number of output bytes. */
ORC_API void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* marching core shared by train / inference / distill (raymarching.cu:359-400, 427-479, 750-804,    */
/* 864-925), with the rounding sequence of SURVEY.md Appendix A.                                      */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, Hf, H3, Cf, half_H;
    uint32_t H;
    const uint8_t* grid;
} march_ctx;

static void march_ctx_init(march_ctx* m, const float* o, const float* d, const uint8_t* grid, float bound,
                           float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1.0f / m->dx; m->rdy = 1.0f / m->dy; m->rdz = 1.0f / m->dz;      /* :337 */
    m->rH = 1.0f / (float)H;                                                   /* :338 */
    m->Hf = (float)H;
    m->H3 = (float)(H * H * H);                                                /* :339 (declared float) */
    m->Cf = (float)C;
    m->half_H = 0.5f * (float)H;
    m->H = H;
    m->bound = bound;
    m->dt_gamma = dt_gamma;
    m->dt_min = TWO_SQRT3 / (float)max_steps;                                  /* :345 */
    m->dt_max = (TWO_SQRT3 * (float)(1u << (C - 1))) / (float)H;               /* :346 */
    m->grid = grid;
}

static inline float march_dt(const march_ctx* m, float t) { return clampf(t * m->dt_gamma, m->dt_min, m->dt_max); }

/* One visit of the loop body at ray parameter t (:361-379).  Returns occupancy; fills the clamped
 * point, cell and level so that the caller can emit or skip.  *index_out is the bitfield bit index. */
static inline int march_probe(const march_ctx* m, float t, float dt, float* x, float* y, float* z, int* nx, int* ny,
                              int* nz, float* mip_bound, uint32_t* index_out) {
    *x = clampf(fmaf(t, m->dx, m->ox), -m->bound, m->bound);
    *y = clampf(fmaf(t, m->dy, m->oy), -m->bound, m->bound);
    *z = clampf(fmaf(t, m->dz, m->oz), -m->bound, m->bound);
    const int l1 = mip_from_pos(*x, *y, *z, m->Cf), l2 = mip_from_dt(dt, m->Hf, m->Cf);
    const int level = l1 > l2 ? l1 : l2;
    *mip_bound = fminf(scalbnf(1.0f, level), m->bound);
    const float mip_rbound = 1.0f / *mip_bound;
    /* 0.5 * (x * mip_rbound + 1) * H: the inner sum contracts to one fma; the double detour rounds once,
     * which equals one float multiply by the exactly-representable 0.5*H.  (int) truncates. */
    *nx = (int)clampf(fmaf(*x, mip_rbound, 1.0f) * m->half_H, 0.0f, (float)(m->H - 1));
    *ny = (int)clampf(fmaf(*y, mip_rbound, 1.0f) * m->half_H, 0.0f, (float)(m->H - 1));
    *nz = (int)clampf(fmaf(*z, mip_rbound, 1.0f) * m->half_H, 0.0f, (float)(m->H - 1));
    /* index is evaluated in float because H3 is a float (:378) */
    const uint32_t index = (uint32_t)fmaf((float)level, m->H3, (float)morton3D_1((uint32_t)*nx, (uint32_t)*ny, (uint32_t)*nz));
    *index_out = index;
    return (m->grid[index >> 3] >> (index & 7)) & 1;
}

/* Skip to the exit of the current voxel (:388-399). */
static inline float march_skip(const march_ctx* m, float t, float x, float y, float z, int nx, int ny, int nz,
                               float mip_bound) {
    const float tx = fmaf(mip_bound, fmaf((fmaf(signf1(m->dx), 0.5f, (float)nx + 0.5f)) * m->rH, 2.0f, -1.0f), -x) * m->rdx;
    const float ty = fmaf(mip_bound, fmaf((fmaf(signf1(m->dy), 0.5f, (float)ny + 0.5f)) * m->rH, 2.0f, -1.0f), -y) * m->rdy;
    const float tz = fmaf(mip_bound, fmaf((fmaf(signf1(m->dz), 0.5f, (float)nz + 0.5f)) * m->rH, 2.0f, -1.0f), -z) * m->rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        t += march_dt(m, t);
    } while (t < tt);
    return t;
}

/* kernel_march_rays_train, :311-480.
 * Deviation made explicit: the reference reserves output slots with atomicAdd, so its (offset, row) assignment
 * is a run-dependent permutation.  The oracle -- like the new CUDA kernel -- uses the canonical
 * assignment: row n of `rays` is ray n, offsets are the exclusive prefix sum of counts in ray-id order, starting
 * from the incoming counter[0].  counter[0] += sum(counts), counter[1] += N, exactly as the atomics total.
 * Rays with offset + count > M write nothing (:416).  Buffers must be zero-initialised by the caller (:205-207). */
ORC_API void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                  float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                  const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                  int32_t* rays, int32_t* counter, const float* noises) {
    uint32_t point_index = (uint32_t)counter[0];
    for (uint32_t n = 0; n < N; n++) {
        march_ctx m;
        march_ctx_init(&m, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n], noise = noises[n];
        const float t0 = fmaf(clampf(near * dt_gamma, m.dt_min, m.dt_max), noise, near); /* :348-351 */
        float x, y, z, mb;
        int nx, ny, nz;
        uint32_t idx;
        /* pass 1 (:354-400) */
        float t = t0;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            const float dt = march_dt(&m, t);
            if (march_probe(&m, t, dt, &x, &y, &z, &nx, &ny, &nz, &mb, &idx)) {
                num_steps++;
                t += dt;
            } else {
                t = march_skip(&m, t, x, y, z, nx, ny, nz, mb);
            }
        }
        rays[n * 3] = (int32_t)n;
        rays[n * 3 + 1] = (int32_t)point_index;
        rays[n * 3 + 2] = (int32_t)num_steps;
        const uint32_t my_off = point_index;
        point_index += num_steps;
        if (num_steps == 0) continue;
        if (my_off + num_steps > M) continue;
        /* pass 2 (:418-479) */
        float* px = xyzs + (size_t)my_off * 3;
        float* pd = dirs + (size_t)my_off * 3;
        float* pl = deltas + (size_t)my_off * 2;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            const float dt = march_dt(&m, t);
            if (march_probe(&m, t, dt, &x, &y, &z, &nx, &ny, &nz, &mb, &idx)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt;
                pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            } else {
                t = march_skip(&m, t, x, y, z, nx, ny, nz, mb);
            }
        }
    }
    counter[0] = (int32_t)point_index;
    counter[1] += (int32_t)N;
}

/* kernel_march_rays :700-805 and kernel_march_rays_distill :811-926 (edit_grid/edit_occ may be NULL).
 * Output buffers must be zero-initialised by the caller (raymarching.py:334-336, 394-397). */
static void march_rays_impl(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                            const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t C, uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* nears,
                            const float* fars, float* xyzs, float* dirs, float* deltas, uint8_t* edit_occ,
                            const float* noises, uint8_t* inexact) {
    for (uint32_t n = 0; n < n_alive; n++) {
        if (inexact) inexact[n] = 0;
        const int32_t index = rays_alive[n];
        const float noise = noises[n];
        march_ctx m;
        march_ctx_init(&m, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma, max_steps, C, H);
        float* px = xyzs + (size_t)n * n_step * 3;
        float* pd = dirs + (size_t)n * n_step * 3;
        float* pl = deltas + (size_t)n * n_step * 2;
        uint8_t* pe = edit_occ ? edit_occ + (size_t)n * n_step : NULL;
        float t = rays_t[index];
        const float far = fars[index];
        (void)nears;
        uint32_t step = 0;
        t = fmaf(clampf(t * dt_gamma, m.dt_min, m.dt_max), noise, t); /* :746 */
        float last_t = t;
        float x, y, z, mb;
        int nx, ny, nz;
        uint32_t idx;
        while (t < far && step < n_step) {
            const float dt = march_dt(&m, t);
            if (march_probe(&m, t, dt, &x, &y, &z, &nx, &ny, &nz, &mb, &idx)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt;
                pl[1] = t - last_t;
                /* test bookkeeping: is the delta the compositor will add back to rays_t (:1006) an EXACT difference?  If
                 * fl(last_t + delta) == t for every sample, rays_t is rebuilt to the marcher's own t wherever a round ends. */
                if (inexact) { volatile float back = last_t + pl[1]; if (back != t) inexact[n] = 1; }
                last_t = t;
                px += 3; pd += 3; pl += 2;
                if (pe) {
                    if ((edit_grid[idx >> 3] >> (idx & 7)) & 1) *pe = 1; /* :885, :906-909 */
                    pe++;
                }
                step++;
            } else {
                t = march_skip(&m, t, x, y, z, nx, ny, nz, mb);
            }
        }
    }
}

ORC_API void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                            const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t C, uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* nears,
                            const float* fars, float* xyzs, float* dirs, float* deltas, uint8_t* edit_occ,
                            const float* noises) {
    march_rays_impl(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, edit_grid, nears, fars, xyzs,
                    dirs, deltas, edit_occ, noises, NULL);
}

/* The same kernel, additionally reporting per alive slot whether any delta emitted in this round was NOT an exact difference
 * (see march_rays_impl).  Not a reference output: tests/test_schedule_theory.py checks the product's exactness criterion with it. */
ORC_API void orc_march_rays_track(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                                  const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                                  uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                                  float* dirs, float* deltas, const float* noises, uint8_t* inexact) {
    march_rays_impl(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, NULL, nears, fars, xyzs, dirs,
                    deltas, NULL, noises, inexact);
}

/* ------------------------------------------------------------------------------------------------ */
/* compositing: raymarching.cu:500-682, 948-1142.  __expf is the device fast exponential; the oracle  */
/* uses expf and parity is tolerance-level (rtol 1e-4).                                 */
/* ------------------------------------------------------------------------------------------------ */

/* kernel_composite_rays_train_forward, :500-577 */
ORC_API void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                              const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                              float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float* ps = sigmas + offset;
        const float* pc = rgbs + (size_t)offset * 3;
        const float* pl = deltas + (size_t)offset * 2;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const float alpha = 1.0f - expf(-ps[0] * pl[0]);
            const float weight = alpha * T;
            r += weight * pc[0]; g += weight * pc[1]; b += weight * pc[2];
            t += pl[1];
            d += weight * t;
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            ps++; pc += 3; pl += 2;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* kernel_composite_rays_train_backward, :601-682.  grad buffers must be zero-initialised (raymarching.py:283-284). */
ORC_API void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                               const float* sigmas, const float* rgbs, const float* deltas,
                                               const int32_t* rays, const float* weights_sum, const float* image,
                                               uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                               float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float* gi = grad_image + (size_t)index * 3;
        const float ws_final = weights_sum[index];
        const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2];
        const float* ps = sigmas + offset;
        const float* pc = rgbs + (size_t)offset * 3;
        const float* pl = deltas + (size_t)offset * 2;
        float* gs = grad_sigmas + offset;
        float* gc = grad_rgbs + (size_t)offset * 3;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const float alpha = 1.0f - expf(-ps[0] * pl[0]);
            const float weight = alpha * T;
            r += weight * pc[0]; g += weight * pc[1]; b += weight * pc[2];
            ws += weight;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * weight; gc[1] = gi[1] * weight; gc[2] = gi[2] * weight;
            gs[0] = pl[0] * (gi[0] * (T * pc[0] - (r_final - r)) + gi[1] * (T * pc[1] - (g_final - g)) +
                             gi[2] * (T * pc[2] - (b_final - b)) + gws * (1 - ws_final));
            if (T < T_thresh) break;
            ps++; pc += 3; pl += 2; gs++; gc += 3;
        }
    }
}

/* kernel_composite_rays :948-1035 and kernel_composite_rays_distill :1037-1142
 * (weights_edit_sum / depth_edit / edit_occ NULL for the plain variant). */
static void composite_rays_impl(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                float* weights_edit_sum, float* depth, float* depth_edit, const uint8_t* edit_occ,
                                float* image, int32_t* steps_done) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float* ps = sigmas + (size_t)n * n_step;
        const float* pc = rgbs + (size_t)n * n_step * 3;
        const float* pl = deltas + (size_t)n * n_step * 2;
        const uint8_t* pe = edit_occ ? edit_occ + (size_t)n * n_step : NULL;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float weight_edit_sum = weights_edit_sum ? weights_edit_sum[index] : 0.0f;
        float d_edit = depth_edit ? depth_edit[index] : 0.0f;
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (pl[0] == 0) break;
            const float alpha = 1.0f - expf(-ps[0] * pl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            /* nvcc contracts every `acc += weight * v` below into one FFMA (checked against the reference build:
             * tests/golden/ref_infer_lego.npz, ref_distill_flower.npz are reproduced bit for bit only with fmaf) */
            if (pe && *pe) { /* :1098-1101: uses t BEFORE the increment */
                weight_edit_sum += weight;
                d_edit = fmaf(weight, t, d_edit);
            }
            t += pl[1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, pc[0], r); g = fmaf(weight, pc[1], g); b = fmaf(weight, pc[2], b);
            if (T < T_thresh) break;
            ps++; pc += 3; pl += 2;
            if (pe) pe++;
            step++;
        }
        if (steps_done) steps_done[n] = (int32_t)step; /* bookkeeping for the tests only: `step` of :1009 when the loop ends */
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        if (weights_edit_sum) weights_edit_sum[index] = weight_edit_sum;
        weights_sum[index] = weight_sum;
        depth[index] = d;
        if (depth_edit) depth_edit[index] = d_edit;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

ORC_API void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                float* weights_edit_sum, float* depth, float* depth_edit, const uint8_t* edit_occ,
                                float* image) {
    composite_rays_impl(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, weights_edit_sum, depth, depth_edit,
                        edit_occ, image, NULL);
}

/* The same kernel, additionally reporting for every alive slot how many samples the ray completed in this round (the value of the
 * reference's loop counter `step` when its while loop ends).  Not a reference output: the tests of the round-schedule theory
 * (tests/test_schedule_theory.py) need to know at which sample a ray dies. */
ORC_API void orc_composite_rays_steps(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                      const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                                      float* image, int32_t* steps_done) {
    composite_rays_impl(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, NULL, depth, NULL, NULL, image,
                        steps_done);
}

/* ------------------------------------------------------------------------------------------------ */
/* hash-grid encoder: gridencoder/src/gridencoder.cu:50-340 (D = 3 only on the hot path, D in 2..3 here) */
/* fp32 oracle: accumulates in float (the reference accumulates in scalar_t; tolerance in the tests).  */
/* ------------------------------------------------------------------------------------------------ */

static const uint32_t PRIMES[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u}; /* :54 */

/* get_grid_index, :66-84 (without the "* C + ch") */
static inline uint32_t grid_index(uint32_t D, uint32_t gridtype, int align_corners, uint32_t hashmap_size,
                                  uint32_t resolution, const uint32_t* pos_grid) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        uint32_t h = 0;
        for (uint32_t d = 0; d < D; d++) h ^= pos_grid[d] * PRIMES[d];
        index = h;
    }
    return index % hashmap_size;
}

/* Per-level scale as the device computes it: exp2f(level * S) * H - 1 (:138), the multiply-subtract contracted.
 * Device exp2f is ex2.approx (2 ulp); libm exp2f may differ in the last bits, so callers that need bit-exact
 * cell indices pass the device-computed scales through `scales` (NULL = compute here). */
ORC_API void orc_grid_level_scales(uint32_t L, float S, uint32_t H, float* scales) {
    for (uint32_t l = 0; l < L; l++) scales[l] = fmaf(exp2f((float)l * S), (float)H, -1.0f);
}

/* kernel_grid, :87-245.  outputs: [L, B, C] (layout 0, the kernel's native) or [B, L*C] (layout 1, what
 * grid.py:57 returns).  dy_dx (optional): [B, L, D, C]. */
ORC_API void orc_grid_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                                     uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float* dy_dx,
                                     uint32_t gridtype, int align_corners, uint32_t interp, const float* scales,
                                     int out_layout) {
    for (uint32_t level = 0; level < L; level++) {
        const float* grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = scales ? scales[level] : fmaf(exp2f((float)level * S), (float)H, -1.0f);
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b * D;
            float* out = out_layout == 0 ? outputs + ((size_t)level * B + b) * C : outputs + ((size_t)b * L + level) * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) {
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dy_dx) memset(dy_dx + ((size_t)b * L + level) * D * C, 0, sizeof(float) * D * C);
                continue;
            }
            float pos[4], pos_deriv[4];
            uint32_t pos_grid[4];
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = fmaf(in[d], scale, align_corners ? 0.0f : 0.5f);
                pos_grid[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pos_grid[d];
                if (interp == 1) {
                    pos_deriv[d] = 6 * pos[d] * (1.0f - pos[d]);
                    pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
                } else {
                    pos_deriv[d] = 1.0f;
                }
            }
            float results[8] = {0};
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pgl[4];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pgl) * C;
                for (uint32_t ch = 0; ch < C; ch++) results[ch] += w * grid[index + ch];
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = results[ch];
            if (dy_dx) {
                float* dd = dy_dx + ((size_t)b * L + level) * D * C;
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[8] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = scale;
                        uint32_t pgl[4];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                            else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                        }
                        pgl[gd] = pos_grid[gd];
                        const uint32_t il = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pgl) * C;
                        pgl[gd] = pos_grid[gd] + 1;
                        const uint32_t ir = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pgl) * C;
                        for (uint32_t ch = 0; ch < C; ch++) rg[ch] += w * (grid[ir + ch] - grid[il + ch]) * pos_deriv[gd];
                    }
                    for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = rg[ch];
                }
            }
        }
    }
}

/* kernel_grid_backward, :248-340 (+ kernel_input_backward :343-369 when dy_dx != NULL).
 * grad: [L, B, C] (layout 0) or [B, L*C] (layout 1).  grad_embeddings must be zero-initialised (grid.py:77).
 * Accumulates in double so that the oracle itself is order-independent; the device sums are atomic (any order). */
ORC_API void orc_grid_encode_backward(const float* grad, const float* inputs, const int32_t* offsets,
                                      double* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                      uint32_t H, const float* dy_dx, float* grad_inputs, uint32_t gridtype,
                                      int align_corners, uint32_t interp, const float* scales, int grad_layout) {
    for (uint32_t level = 0; level < L; level++) {
        double* gg = grad_embeddings + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = scales ? scales[level] : fmaf(exp2f((float)level * S), (float)H, -1.0f);
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b * D;
            const float* g = grad_layout == 0 ? grad + ((size_t)level * B + b) * C : grad + ((size_t)b * L + level) * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            float pos[4];
            uint32_t pos_grid[4];
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = fmaf(in[d], scale, align_corners ? 0.0f : 0.5f);
                pos_grid[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pos_grid[d];
                if (interp == 1) pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
            }
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pgl[4];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                    else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
                }
                const uint32_t index = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pgl) * C;
                for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += (double)(w * g[ch]);
            }
        }
    }
    if (dy_dx && grad_inputs) {
        for (uint32_t b = 0; b < B; b++)
            for (uint32_t d = 0; d < D; d++) {
                float result = 0;
                for (uint32_t l = 0; l < L; l++)
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float g = grad_layout == 0 ? grad[((size_t)l * B + b) * C + ch] : grad[((size_t)b * L + l) * C + ch];
                        result += g * dy_dx[(((size_t)b * L + l) * D + d) * C + ch];
                    }
                grad_inputs[(size_t)b * D + d] = result;
            }
    }
}

/* kernel_grad_tv, :506-610.  grad is accumulated in place (float, like the device atomics). */
ORC_API void orc_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int32_t* offsets,
                                      float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      uint32_t gridtype, int align_corners, const float* scales) {
    for (uint32_t level = 0; level < L; level++) {
        const float* grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        float* gg = grad + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const float scale = scales ? scales[level] : fmaf(exp2f((float)level * S), (float)H, -1.0f);
        const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            uint32_t pos_grid[4];
            for (uint32_t d = 0; d < D; d++) pos_grid[d] = (uint32_t)floorf(fmaf(in[d], scale, align_corners ? 0.0f : 0.5f));
            float results[8] = {0}, idelta[8] = {0};
            const uint32_t index = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pos_grid) * C;
            const float w = weight / (2 * D);
            for (uint32_t d = 0; d < D; d++) {
                const uint32_t cur_d = pos_grid[d];
                if (cur_d < resolution) {
                    pos_grid[d] = cur_d + 1;
                    const uint32_t ir = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pos_grid) * C;
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[ir + ch];
                        results[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                if (cur_d > 0) {
                    pos_grid[d] = cur_d - 1;
                    const uint32_t il = grid_index(D, gridtype, align_corners, hashmap_size, resolution, pos_grid) * C;
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[il + ch];
                        results[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                pos_grid[d] = cur_d;
            }
            for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += w * results[ch] * (1.0f / sqrtf(idelta[ch] + 1e-9f));
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* fully fused MLP: ffmlp/src/ffmlp.cu:331-407 (fwd), :410-518 + :749-895 (bwd); layout ffmlp.cu:631-634, */
/* ffmlp/ffmlp.py:121.  y = W_last . relu(... relu(W0 . x)), no bias, no output activation.           */
/* ------------------------------------------------------------------------------------------------ */

/* round-to-nearest-even float -> IEEE half -> float, emulating the fp16 storage of activations */
static inline float round_half(float f) {
    _Float16 h = (_Float16)f;
    return (float)h;
}

static inline float act_fwd(uint32_t activation, float x) { /* utils.h:424-475 */
    switch (activation) {
        case 0: return x > 0.0f ? x : 0.0f;
        case 1: return expf(x);
        case 2: return sinf(x);
        case 3: return 1.0f / (1.0f + expf(-x));
        case 4: { const float y = x * 10.0f; return 0.5f * (y + sqrtf(y * y + 4)) / 10.0f; }
        case 5: return logf(expf(x * 10.0f) + 1.0f) / 10.0f;
        default: return x;
    }
}
/* derivative expressed through the stored post-activation value, utils.h:538-583 */
static inline float act_bwd(uint32_t activation, float g, float fwd) {
    switch (activation) {
        case 0: return fwd > 0.0f ? g : 0.0f;
        case 1: return g * fwd;
        case 3: return g * (fwd * (1.0f - fwd));
        case 4: { const float y = fwd * 10.0f; return g * (y * y / (y * y + 1)); }
        case 5: return g * (1.0f - expf(-fwd * 10.0f));
        default: return g;
    }
}

/* inputs [B, in] ; weights flat: [hidden,in] + (num_layers-1) x [hidden,hidden] + [out,hidden] ; outputs [B, out];
 * forward_buffer (optional) [num_layers, B, hidden].  Values are float but rounded to fp16 wherever the device
 * stores fp16 (inputs/weights are expected to be fp16-representable already); accumulation is fp32. */
ORC_API void orc_ffmlp_forward(const float* inputs, const float* weights, uint32_t B, uint32_t input_dim,
                               uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                               uint32_t output_activation, float* forward_buffer, float* outputs) {
    float* cur = (float*)malloc(sizeof(float) * hidden_dim);
    float* nxt = (float*)malloc(sizeof(float) * hidden_dim);
    for (uint32_t b = 0; b < B; b++) {
        const float* x = inputs + (size_t)b * input_dim;
        const float* W = weights;
        for (uint32_t j = 0; j < hidden_dim; j++) {
            float acc = 0;
            for (uint32_t k = 0; k < input_dim; k++) acc += W[(size_t)j * input_dim + k] * x[k];
            cur[j] = round_half(act_fwd(activation, acc));
        }
        if (forward_buffer) memcpy(forward_buffer + ((size_t)0 * B + b) * hidden_dim, cur, sizeof(float) * hidden_dim);
        W += (size_t)hidden_dim * input_dim;
        for (uint32_t l = 1; l < num_layers; l++) {
            for (uint32_t j = 0; j < hidden_dim; j++) {
                float acc = 0;
                for (uint32_t k = 0; k < hidden_dim; k++) acc += W[(size_t)j * hidden_dim + k] * cur[k];
                nxt[j] = round_half(act_fwd(activation, acc));
            }
            float* tmp = cur; cur = nxt; nxt = tmp;
            if (forward_buffer) memcpy(forward_buffer + ((size_t)l * B + b) * hidden_dim, cur, sizeof(float) * hidden_dim);
            W += (size_t)hidden_dim * hidden_dim;
        }
        for (uint32_t j = 0; j < output_dim; j++) {
            float acc = 0;
            for (uint32_t k = 0; k < hidden_dim; k++) acc += W[(size_t)j * hidden_dim + k] * cur[k];
            outputs[(size_t)b * output_dim + j] = round_half(act_fwd(output_activation, acc));
        }
    }
    free(cur); free(nxt);
}

/* grad [B,out]; forward_buffer [num_layers,B,hidden] from the forward; grad_weights flat like weights (double,
 * zero-initialised by the caller); grad_inputs [B,in] or NULL.  Hidden-gradient activations are rounded to fp16
 * between layers as the device stores them (backward_buffer, ffmlp.py:73). */
ORC_API void orc_ffmlp_backward(const float* grad, const float* inputs, const float* weights, const float* forward_buffer,
                                uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                uint32_t num_layers, uint32_t activation, double* grad_weights, float* grad_inputs) {
    const size_t w_first = (size_t)hidden_dim * input_dim, w_hid = (size_t)hidden_dim * hidden_dim;
    const float* W_last = weights + w_first + (num_layers - 1) * w_hid;
    double* gW_last = grad_weights + w_first + (num_layers - 1) * w_hid;
    float* dcur = (float*)malloc(sizeof(float) * hidden_dim);
    float* dnxt = (float*)malloc(sizeof(float) * hidden_dim);
    for (uint32_t b = 0; b < B; b++) {
        const float* g = grad + (size_t)b * output_dim;
        const float* h_last = forward_buffer + ((size_t)(num_layers - 1) * B + b) * hidden_dim;
        /* dW_last = g^T h_last (ffmlp.cu:804-810) ; dh = (W_last^T g) * act'(h_last) (:452-486) */
        for (uint32_t j = 0; j < output_dim; j++)
            for (uint32_t k = 0; k < hidden_dim; k++) gW_last[(size_t)j * hidden_dim + k] += (double)(g[j] * h_last[k]);
        for (uint32_t k = 0; k < hidden_dim; k++) {
            float acc = 0;
            for (uint32_t j = 0; j < output_dim; j++) acc += W_last[(size_t)j * hidden_dim + k] * g[j];
            dcur[k] = round_half(act_bwd(activation, acc, h_last[k]));
        }
        /* hidden layers, last to first (:508-510, :844-863) */
        for (uint32_t l = num_layers - 1; l >= 1; l--) {
            const float* W = weights + w_first + (size_t)(l - 1) * w_hid;  /* maps h_{l-1} -> h_l */
            double* gW = grad_weights + w_first + (size_t)(l - 1) * w_hid;
            const float* h_prev = forward_buffer + ((size_t)(l - 1) * B + b) * hidden_dim;
            for (uint32_t j = 0; j < hidden_dim; j++)
                for (uint32_t k = 0; k < hidden_dim; k++) gW[(size_t)j * hidden_dim + k] += (double)(dcur[j] * h_prev[k]);
            for (uint32_t k = 0; k < hidden_dim; k++) {
                float acc = 0;
                for (uint32_t j = 0; j < hidden_dim; j++) acc += W[(size_t)j * hidden_dim + k] * dcur[j];
                dnxt[k] = round_half(act_bwd(activation, acc, h_prev[k]));
            }
            float* tmp = dcur; dcur = dnxt; dnxt = tmp;
        }
        /* input layer (:866-887) */
        const float* x = inputs + (size_t)b * input_dim;
        for (uint32_t j = 0; j < hidden_dim; j++)
            for (uint32_t k = 0; k < input_dim; k++) grad_weights[(size_t)j * input_dim + k] += (double)(dcur[j] * x[k]);
        if (grad_inputs) {
            for (uint32_t k = 0; k < input_dim; k++) {
                float acc = 0;
                for (uint32_t j = 0; j < hidden_dim; j++) acc += weights[(size_t)j * input_dim + k] * dcur[j];
                grad_inputs[(size_t)b * input_dim + k] = round_half(acc);
            }
        }
    }
    free(dcur); free(dnxt);
}

/* ------------------------------------------------------------------------------------------------ */
/* spherical harmonics direction encoding (adjacent row f-1): shencoder/src/shencoder.cu:27-123       */
/* degree <= 4 (the NeRF colour net uses 4 -> 16 channels, LAENeRF's offset net 3 -> 9).               */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_sh_encode(const float* dirs, uint32_t B, uint32_t degree, float* outputs) {
    const uint32_t C2 = degree * degree;
    for (uint32_t b = 0; b < B; b++) {
        const float x = dirs[b * 3], y = dirs[b * 3 + 1], z = dirs[b * 3 + 2];
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        float* o = outputs + (size_t)b * C2;
        float v[16];
        v[0] = 0.28209479177387814f;
        v[1] = -0.48860251190291987f * y; v[2] = 0.48860251190291987f * z; v[3] = -0.48860251190291987f * x;
        v[4] = 1.0925484305920792f * xy; v[5] = -1.0925484305920792f * yz;
        v[6] = 0.94617469575755997f * z2 - 0.31539156525251999f; v[7] = -1.0925484305920792f * xz;
        v[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
        v[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2); v[10] = 2.8906114426405538f * xy * z;
        v[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2); v[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
        v[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2); v[14] = 1.4453057213202769f * z * (x2 - y2);
        v[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
        for (uint32_t c = 0; c < C2 && c < 16; c++) o[c] = v[c];
    }
}
