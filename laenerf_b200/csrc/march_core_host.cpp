// march_core_host.cpp -- host build of march_core.cuh for CPU unit tests (tests/test_march_core.py).
//
// Two things are checked without a GPU:
//   1. march_window<G> (closed-form window of the t-sequence) against plain sequential float adds;
//   2. the lane-group resolve algorithm of raymarch.cu (march_group), emulated lane by lane with the same
//      march_core arithmetic, against the sequential oracle -- i.e. that "evaluate a window in parallel, then
//      resolve visits with ballots" reproduces the reference's visit order exactly.
// Not part of liblaenerf_b200.so.
#include <stdint.h>
#include <math.h>
#include <vector>
#include "march_core.cuh"

using namespace lnrf;

template <int G>
static uint32_t emulate_group(const MarchParams& p, const Ray& r, const uint8_t* grid, float t, float far, uint32_t max_emit,
                              float* tl) {
    uint32_t cnt = 0;
    float pend = -INFINITY;
    bool alive = true;
    while (alive) {
        float s[G], nxt = 0.f;
        Probe q[G];
        bool valid[G], occ[G];
        for (int l = 0; l < G; l++) {
            s[l] = march_window<G>(p, t, l, &nxt);
            const float dt = p.dt_const ? p.dt0 : march_dt(p, s[l]);
            valid[l] = s[l] < far;
            occ[l] = false;
            q[l].tt = 0.f;
            if (valid[l]) {
                q[l] = march_probe(p, r, s[l], dt);
                occ[l] = (grid[q[l].index >> 3] >> (q[l].index & 7u)) & 1u;
            }
        }
        int v = G;
        for (int l = 0; l < G; l++)
            if (s[l] >= pend) { v = l; break; }
        if (v < G) pend = -INFINITY;
        bool res = v < G;
        bool vis[G];
        for (int l = 0; l < G; l++) vis[l] = false;
        uint32_t nvis = 0;
        while (res) {
            if (!valid[v]) { alive = false; res = false; }
            else if (occ[v]) {
                int run = 0;
                while (v + run < G && occ[v + run]) run++;
                const int room = (int)(max_emit - cnt) - (int)nvis;
                if (run >= room) { run = room; alive = false; res = false; }
                for (int k = 0; k < run; k++) vis[v + k] = true;
                nvis += run;
                v += run;
                if (v >= G) res = false;
            } else {
                const float tt = q[v].tt;
                int j = G;
                for (int l = v + 1; l < G; l++)
                    if (s[l] >= tt) { j = l; break; }
                if (j < G) v = j;
                else { pend = tt; v = G; res = false; }
            }
        }
        for (int l = 0; l < G; l++)
            if (vis[l]) tl[cnt++] = s[l];
        t = nxt;
    }
    return cnt;
}

extern "C" {

// returns the number of mismatching members over `windows` consecutive windows starting at t
uint64_t mch_check_window(float t, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, int G, uint32_t windows) {
    const MarchParams p = make_march_params(1.0f, dt_gamma, max_steps, C, H);
    uint64_t bad = 0;
    for (uint32_t w = 0; w < windows; w++) {
        float seq[33];
        seq[0] = t;
        for (int k = 1; k <= G; k++) seq[k] = seq[k - 1] + march_dt(p, seq[k - 1]);
        float nxt = 0.f;
        for (int l = 0; l < G; l++) {
            float n2;
            const float m = (G == 32) ? march_window<32>(p, t, l, &n2) : march_window<8>(p, t, l, &n2);
            if (m != seq[l]) bad++;
            nxt = n2;
        }
        if (nxt != seq[G]) bad++;
        t = nxt;
    }
    return bad;
}

// Emulated group march of N rays; counts[n] and, concatenated in ray order, the visited t values (ts, capacity cap).
// Returns the total number of samples.
uint64_t mch_group_march(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars,
                         const float* noises, int G, uint32_t* counts, float* ts, uint64_t cap) {
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
    std::vector<float> tl(max_steps);
    uint64_t total = 0;
    for (uint32_t n = 0; n < N; n++) {
        const Ray r = make_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        const float near = nears[n];
        const float t0 = f_fma(f_clamp(f_mul(near, p.dt_gamma), p.dt_min, p.dt_max), noises[n], near);
        const uint32_t c = (G == 32) ? emulate_group<32>(p, r, grid, t0, fars[n], max_steps, tl.data())
                                     : emulate_group<8>(p, r, grid, t0, fars[n], max_steps, tl.data());
        counts[n] = c;
        for (uint32_t k = 0; k < c && total + k < cap; k++) ts[total + k] = tl[k];
        total += c;
    }
    return total;
}
}
