// march_core.cuh -- the occupancy-grid marching arithmetic, written once for device and host.
//
// Restates the per-visit arithmetic of the reference marcher (raymarching/src/raymarching.cu:335-400, shared by
// the inference kernels :730-803 and the distill kernels :844-925) with every rounding step pinned by an explicit
// intrinsic so that results are bit-identical to the reference build (the contraction pattern nvcc applies to the
// reference source is listed in SURVEY.md Appendix A).  Compiles for the host as well (LNRF_HD) so the closed-form
// window generator below can be checked against plain sequential float adds without a GPU (tests/test_march_core.py).
//
// Key observation the B200 design rests on: whichever branch the reference takes (occupied: t += dt; empty: the
// do/while skip loop), t only ever advances by  t <- t + clamp(t*dt_gamma, dt_min, dt_max).  Every t the reference
// visits is therefore a member of ONE sequence S(t0) that does not depend on the occupancy grid; the grid only
// selects which members are visited.  A group of lanes can evaluate a window of consecutive members of S in
// parallel (cell lookup, skip target) and then resolve "which of them does the reference visit" with ballots.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define LNRF_HD __host__ __device__ __forceinline__
#else
#define LNRF_HD inline
#endif

namespace lnrf {

#if defined(__CUDA_ARCH__)
LNRF_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
LNRF_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
LNRF_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
LNRF_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
LNRF_HD uint32_t f2u(float f) { return __float_as_uint(f); }
LNRF_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
// host build uses -ffp-contract=off, so these are single correctly-rounded operations
LNRF_HD float f_mul(float a, float b) { return a * b; }
LNRF_HD float f_add(float a, float b) { return a + b; }
LNRF_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
LNRF_HD float f_div(float a, float b) { return a / b; }
LNRF_HD uint32_t f2u(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
LNRF_HD float u2f(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }
#endif

LNRF_HD float f_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// frexpf exponent: [0.5,1) -> 0, [1,2) -> 1, ...; 0 -> 0; denormals handled like libm/CUDA.
LNRF_HD int frexp_exponent(float ax /* >= 0 */) {
    uint32_t b = f2u(ax);
    if (b == 0u) return 0;
    int e = (int)(b >> 23);
    if (e == 0) {  // denormal
        b = f2u(f_mul(ax, 16777216.0f));
        return (int)(b >> 23) - 126 - 24;
    }
    if (e == 255) return 0;  // inf/nan: frexp leaves the exponent unspecified; never reached on clamped points
    return e - 126;
}

LNRF_HD uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
LNRF_HD uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits10(x) | (expand_bits10(y) << 1) | (expand_bits10(z) << 2);
}
LNRF_HD uint32_t morton3d_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// Per-launch constants (identical for every ray).
struct MarchParams {
    float bound, neg_bound, rbound, dt_gamma, dt_min, dt_max, rH, Hf, H3, Cm1, half_H, Hm1;
    uint32_t max_steps;
    int dt_const;  // dt_gamma == 0 (every shipped LAENeRF config): dt is the same for every t ...
    float dt0;     // ... namely clamp(0, dt_min, dt_max) (== dt_min unless max_steps is so small that dt_min > dt_max)
    int jump;      // closed-form windows are resolved with the jump table (march_jump + pointer doubling) instead of the serial loop
    int fast_forward;  // a pending skip moves the group straight to the member it lands on (march_fast_forward)
};

LNRF_HD MarchParams make_march_params(float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    MarchParams p;
    p.bound = bound;
    p.neg_bound = -bound;
    p.rbound = f_div(1.0f, bound);  // the reference divides per visit (1 / mip_bound); the quotient only takes these values
    p.dt_gamma = dt_gamma;
    p.dt_min = f_div(3.4641015529632568f, (float)max_steps);                      // 2*SQRT3()/max_steps
    p.dt_max = f_div(f_mul(3.4641015529632568f, (float)(1u << (C - 1))), (float)H);  // 2*SQRT3()*(1<<(C-1))/H
    p.rH = f_div(1.0f, (float)H);
    p.Hf = (float)H;
    p.H3 = (float)(H * H * H);
    p.Cm1 = (float)C - 1.0f;
    p.half_H = 0.5f * (float)H;
    p.Hm1 = (float)(H - 1);
    p.max_steps = max_steps;
    p.dt_const = (dt_gamma == 0.0f) ? 1 : 0;
    p.dt0 = f_clamp(0.0f, p.dt_min, p.dt_max);
    p.jump = p.dt_const;
    p.fast_forward = p.dt_const && C > 1;  // a cascade-0 voxel spans < 5 members: nothing to skip over
    return p;
}

LNRF_HD float march_dt(const MarchParams& p, float t) { return f_clamp(f_mul(t, p.dt_gamma), p.dt_min, p.dt_max); }

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, sx, sy, sz;  // s* = 0.5 * sign(d*)
};

LNRF_HD Ray make_ray(const float* o, const float* d) {
    Ray r;
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.rdx = f_div(1.0f, r.dx); r.rdy = f_div(1.0f, r.dy); r.rdz = f_div(1.0f, r.dz);
    r.sx = copysignf(0.5f, r.dx); r.sy = copysignf(0.5f, r.dy); r.sz = copysignf(0.5f, r.dz);
    return r;
}

// End of the useful part of a ray: min(far, where it leaves `box` = {lo xyz, hi xyz}), a box that contains every occupied cell with
// a margin (lnrf_render_desc.occupied_box).  Beyond it the reference's walk only skips empty cells, so stopping there emits the same
// samples; a ray that misses the box emits none (-inf).  The rounding of this slab test (a few ulp of t) is far inside the margin.
// null box: far unchanged.
LNRF_HD float clip_far_to_box(const float* box, const Ray& r, float far) {
    if (!box) return far;
    const float o[3] = {r.ox, r.oy, r.oz}, rd[3] = {r.rdx, r.rdy, r.rdz};
    float t_in = -INFINITY, t_out = INFINITY;
    for (int a = 0; a < 3; a++) {
        const float t1 = (box[a] - o[a]) * rd[a], t2 = (box[3 + a] - o[a]) * rd[a];  // +-inf for d = 0, NaN on a face: fmin / fmax drop NaN
        t_in = fmaxf(t_in, fminf(t1, t2));
        t_out = fminf(t_out, fmaxf(t1, t2));
    }
    if (!(t_in <= t_out)) return -INFINITY;
    return fminf(far, t_out);
}

struct Probe {
    float x, y, z;    // clamped sample position
    float tt;         // where the reference's skip loop must get to if this cell is empty
    uint32_t index;   // bit index into the occupancy bitfield
};

// Everything the reference computes at one visited t except the bitfield load itself.
LNRF_HD Probe march_probe(const MarchParams& p, const Ray& r, float t, float dt) {
    Probe q;
    q.x = f_clamp(f_fma(t, r.dx, r.ox), p.neg_bound, p.bound);
    q.y = f_clamp(f_fma(t, r.dy, r.oy), p.neg_bound, p.bound);
    q.z = f_clamp(f_fma(t, r.dz, r.oz), p.neg_bound, p.bound);
    // cascade level (raymarching.cu:42-54).  With a single cascade (bound <= 1: every synthetic scene) both frexp terms clamp
    // to 0, so the branch -- uniform over the launch -- skips them.  1 / mip_bound is exact without a division: mip_bound is
    // 2^level (reciprocal = the power of two with the negated exponent) unless the box itself is smaller (then 1 / bound).
    int level = 0;
    if (p.Cm1 > 0.0f) {
        const float mx = fmaxf(fabsf(q.x), fmaxf(fabsf(q.y), fabsf(q.z)));
        const int l1 = (int)fminf(p.Cm1, fmaxf(0.0f, (float)frexp_exponent(mx)));
        const int l2 = (int)fminf(p.Cm1, fmaxf(0.0f, (float)frexp_exponent(f_mul(f_mul(dt, p.Hf), 0.5f))));
        level = l1 > l2 ? l1 : l2;
    }
    const float pow2 = u2f((uint32_t)(127 + level) << 23);
    const bool boxed = p.bound < pow2;
    const float mip_bound = boxed ? p.bound : pow2;
    const float mip_rbound = boxed ? p.rbound : u2f((uint32_t)(127 - level) << 23);
    const int nx = (int)f_clamp(f_mul(f_fma(q.x, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
    const int ny = (int)f_clamp(f_mul(f_fma(q.y, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
    const int nz = (int)f_clamp(f_mul(f_fma(q.z, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
    q.index = (uint32_t)f_fma((float)level, p.H3, (float)morton3d((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    // exit of this voxel along the ray (raymarching.cu:390-394)
    const float tx = f_mul(f_fma(mip_bound, f_fma(f_mul(f_add(f_add((float)nx, 0.5f), r.sx), p.rH), 2.0f, -1.0f), -q.x), r.rdx);
    const float ty = f_mul(f_fma(mip_bound, f_fma(f_mul(f_add(f_add((float)ny, 0.5f), r.sy), p.rH), 2.0f, -1.0f), -q.y), r.rdy);
    const float tz = f_mul(f_fma(mip_bound, f_fma(f_mul(f_add(f_add((float)nz, 0.5f), r.sz), p.rH), 2.0f, -1.0f), -q.z), r.rdz);
    q.tt = f_add(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return q;
}

// ---------------------------------------------------------------------------------------------------------
// Window generator: members k = 0..G of S starting at t (s_0 = t, s_{k+1} = s_k + dt(s_k)); returns s_k for
// k = `lane` and writes s_G to *next.
//
// With dt constant the recurrence is a chain of G dependent float adds.  Inside one binade every float is a
// multiple of ulp u, and fl(m*u + dt) = (m + q)*u with q = rint(dt/u) as long as dt/u is not an exact tie and the
// sum stays inside the binade.  Then s_k = t + k*q*u EXACTLY, which each lane evaluates with one fma.  Anything
// else (binade crossing, ties, tiny or non-positive t) takes the sequential path, which is the definition.
// ---------------------------------------------------------------------------------------------------------
// What a closed-form window knows about itself: every member is (mt + k*qi) * u inside binade `e` (u = ulp of the binade).
struct WindowInfo {
    int closed;      // 0: the sequential path produced this window (march_jump is not applicable)
    uint32_t e;      // biased exponent of every member
    uint32_t mt;     // mantissa field of the first member
    uint32_t qi;     // step in ulps (1 <= qi <= 2^18)
    float rq;        // ~ 1 / qi
};

template <int G>
LNRF_HD float march_window(const MarchParams& p, float t, int lane, float* next, WindowInfo* wi = nullptr) {
    if (wi) wi->closed = 0;
    if (p.dt_const) {
        const float dt = p.dt0;
        const uint32_t bt = f2u(t);
        const uint32_t e = bt >> 23;  // sign must be 0 and exponent in a sane range
        if (e >= 64u && e <= 200u) {
            const float u = u2f((e - 23u) << 23);          // ulp of t's binade
            const float r = f_mul(dt, u2f((277u - e) << 23));  // dt / u, exact power-of-two scaling
            const float q = rintf(r);
            if (r < 262144.0f /* G*q stays an exact float for G <= 32 */ && q >= 1.0f && fabsf(f_add(r, -q)) != 0.5f) {
                const float end = f_fma(f_mul((float)G, q), u, t);
                if ((f2u(end) >> 23) == e) {
                    *next = end;
                    if (wi) {
                        wi->closed = 1;
                        wi->e = e;
                        wi->mt = bt & 0x7fffffu;
                        wi->qi = (uint32_t)q;
                        wi->rq = f_div(1.0f, q);
                    }
                    return f_fma(f_mul((float)lane, q), u, t);
                }
            }
        }
        float s = t, mine = t;
#pragma unroll
        for (int k = 1; k <= G; k++) {
            s = f_add(s, dt);
            if (k == lane) mine = s;
        }
        *next = s;
        return mine;
    } else {
        float s = t, mine = t;
#pragma unroll
        for (int k = 1; k <= G; k++) {
            s = f_add(s, march_dt(p, s));
            if (k == lane) mine = s;
        }
        *next = s;
        return mine;
    }
}

// Fast-forward over members nobody visits.  When a skip target `pend` lies beyond the current position -- an empty voxel of an
// outer cascade spans dozens of members (bound 16: 0.125 / dt = 37), so a skip used to cost one idle window after the other --
// the group moves straight to the first member >= pend: inside a binade member k is (mt + k*qi) ulps, so that member is
// k = ceil((m_pend - mt) / qi), capped at the last member of the binade (the crossing itself stays with the sequential adds of
// march_window).  Same lattice argument as the closed-form window: t + k*qi*u is EXACTLY the k-th sequential sum.  Returns t
// unchanged when nothing is pending, when the target is within the next G members anyway, or when the closed form does not apply.
LNRF_HD float march_fast_forward(const MarchParams& p, float t, float pend, int G) {
    // cheap filter first (returning t is always valid): only a target at least two windows ahead is worth the integer divisions
    if (!p.dt_const || !(pend > f_fma((float)(2 * G), p.dt0, t))) return t;
    const uint32_t bt = f2u(t);
    const uint32_t e = bt >> 23;
    if (e < 64u || e > 200u) return t;
    const float u = u2f((e - 23u) << 23);
    const float r = f_mul(p.dt0, u2f((277u - e) << 23));
    const float q = rintf(r);
    if (!(r < 262144.0f && q >= 1.0f && fabsf(f_add(r, -q)) != 0.5f)) return t;
    const uint32_t qi = (uint32_t)q, mt = bt & 0x7fffffu;
    const uint32_t kmax = (0x7fffffu - mt) / qi;  // members 0..kmax stay inside the binade
    const uint32_t bp = f2u(pend);
    const uint32_t ep = bp >> 23;                 // pend > t > 0: sign clear; inf / NaN have ep = 255
    uint32_t k = kmax;
    if (ep == e) {
        const uint32_t a = (bp & 0x7fffffu) - mt;
        const uint32_t kk = (a + qi - 1u) / qi;
        k = kk < kmax ? kk : kmax;
    }
    if (k < (uint32_t)G) return t;
    return f_fma((float)(k * qi), u, t);
}

// Successor of member v of a closed-form window whose cell is empty: the reference runs `do { t += dt; } while (t < tt)`
// (raymarching.cu:396-398), i.e. it visits the first member j > v with s_j >= tt.  All members are (mt + j*qi) ulps of one
// binade, so the comparison is integer arithmetic on the mantissa fields: j = ceil((mtt - mt) / qi).  Returns G when no member
// of this window qualifies (tt in a later binade, or past the last member).  The float quotient is only a first guess; the
// result is fixed up with exact integer products, so host and device agree whatever the rounding of rq.
LNRF_HD int march_jump(const WindowInfo& w, int v, float tt, int G) {
    const uint32_t btt = f2u(tt);
    const uint32_t et = btt >> 23;  // tt >= s_v > 0: sign bit clear
    if (et > w.e) return G;
    if (et < w.e) return v + 1;     // not reachable for tt >= s_v; keeps the function total
    const uint32_t mtt = btt & 0x7fffffu;
    if (mtt <= w.mt) return v + 1;
    const uint32_t a = mtt - w.mt;
    const float f = f_mul((float)a, w.rq);
    if (f > 40.0f) return G;
    uint32_t j = (uint32_t)f;
    if (j * w.qi < a) j++;
    if (j * w.qi < a) j++;
    if (j > 0u && (j - 1u) * w.qi >= a) j--;
    const int ji = (int)j;
    return ji <= v ? v + 1 : (ji < G ? ji : G);
}

}  // namespace lnrf
