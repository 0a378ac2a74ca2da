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

static int g_fast_forward = 1;  // mch_set_fast_forward(0): emulate without the skip over unvisited members

template <int G>
static uint32_t emulate_group(const MarchParams& p, const Ray& r, const uint8_t* grid, float t, float far, uint32_t max_emit,
                              float* tl) {
    uint32_t cnt = 0;
    float pend = -INFINITY;
    bool alive = true;
    while (alive) {
        if (g_fast_forward) t = march_fast_forward(p, t, pend, G);
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

// The jump-table resolve of raymarch.cu (march_group, closed-form windows): every lane computes its successor with
// march_jump, the orbit of the first visited lane is found by pointer doubling (log2 G rounds of "shuffles"), and far / sample
// budget / pending-skip bookkeeping is applied to the resulting masks.  Windows that are not closed-form take the serial
// resolve above.  Emulated lane by lane with exactly the device's data flow.
template <int G>
static uint32_t emulate_group_jump(const MarchParams& p, const Ray& r, const uint8_t* grid, float t, float far, uint32_t max_emit,
                                   float* tl, uint64_t* jump_windows) {
    uint32_t cnt = 0;
    float pend = -INFINITY;
    bool alive = true;
    constexpr int LOG = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : 2;
    while (alive) {
        if (g_fast_forward) t = march_fast_forward(p, t, pend, G);
        float s[G], nxt = 0.f;
        Probe q[G];
        bool valid[G], occ[G];
        WindowInfo wi;
        for (int l = 0; l < G; l++) {
            s[l] = march_window<G>(p, t, l, &nxt, &wi);
            const float dt = p.dt_const ? p.dt0 : march_dt(p, s[l]);
            valid[l] = s[l] < far;
            occ[l] = false;
            q[l].tt = 0.f;
            if (valid[l]) {
                q[l] = march_probe(p, r, s[l], dt);
                occ[l] = (grid[q[l].index >> 3] >> (q[l].index & 7u)) & 1u;
            }
        }
        int v0 = G;
        for (int l = 0; l < G; l++)
            if (s[l] >= pend) { v0 = l; break; }
        uint32_t vis = 0;
        if (wi.closed) {
            if (jump_windows) (*jump_windows)++;
            uint32_t occm = 0, valm = 0;
            int nx[G];
            uint32_t orb[G];
            for (int l = 0; l < G; l++) {
                occm |= (uint32_t)occ[l] << l;
                valm |= (uint32_t)valid[l] << l;
                nx[l] = !valid[l] ? G : (occ[l] ? l + 1 : march_jump(wi, l, q[l].tt, G));
                orb[l] = 1u << l;
            }
            for (int rd = 0; rd < LOG; rd++) {
                int n2[G];
                uint32_t o2[G];
                for (int l = 0; l < G; l++) {
                    const int src = nx[l] < G ? nx[l] : G - 1;
                    n2[l] = nx[src];
                    o2[l] = orb[src];
                }
                for (int l = 0; l < G; l++)
                    if (nx[l] < G) { orb[l] |= o2[l]; nx[l] = n2[l]; }
            }
            if (v0 < G) {
                pend = -INFINITY;
                const uint32_t visited = orb[v0];
                const uint32_t inval = visited & ~valm, visv = visited & valm;
                uint32_t emitm = visv & occm;
                const uint32_t room = max_emit - cnt;
                uint32_t ne = 0;
                for (int l = 0; l < G; l++) ne += (emitm >> l) & 1u;
                if (emitm != 0u && ne >= room) {
                    uint32_t keep = 0, k = 0;
                    for (int l = 0; l < G; l++)
                        if ((emitm >> l) & 1u) { if (k < room) keep |= 1u << l; k++; }
                    emitm = keep;
                    alive = false;
                } else if (inval) {
                    alive = false;
                } else {
                    int last = 0;
                    for (int l = 0; l < G; l++)
                        if ((visv >> l) & 1u) last = l;
                    if (!((occm >> last) & 1u)) pend = q[last].tt;
                }
                vis = emitm;
            }
        } else {
            int v = v0;
            if (v < G) pend = -INFINITY;
            bool res = v < G;
            uint32_t nvis = 0;
            while (res) {
                if (!valid[v]) { alive = false; res = false; }
                else if (occ[v]) {
                    int run = 0;
                    while (v + run < G && occ[v + run]) run++;
                    const int room = (int)(max_emit - cnt) - (int)nvis;
                    if (run >= room) { run = room; alive = false; res = false; }
                    for (int k = 0; k < run; k++) vis |= 1u << (v + k);
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
        }
        for (int l = 0; l < G; l++)
            if ((vis >> l) & 1u) tl[cnt++] = s[l];
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

static uint64_t g_jump_windows = 0;
uint64_t mch_jump_windows() { return g_jump_windows; }  // how many windows the jump-table resolve handled so far
void mch_set_fast_forward(int on) { g_fast_forward = on; }
static uint32_t g_max_emit = 0;
void mch_set_max_emit(uint32_t m) { g_max_emit = m; }   // 0: max_steps (training); else the per-call sample budget (inference rounds)

// march_jump against a linear search: over `windows` closed-form windows starting at t, every lane v and the probe targets
// tt = s_v + f * dt for a sweep of f (0 .. 40 steps, including exact member values and half-way points).  Returns mismatches.
uint64_t mch_check_jump(float t, uint32_t max_steps, uint32_t C, uint32_t H, int G, uint32_t windows) {
    const MarchParams p = make_march_params(1.0f, 0.0f, max_steps, C, H);
    uint64_t bad = 0;
    for (uint32_t w = 0; w < windows; w++) {
        float s[33], nxt = 0.f;
        WindowInfo wi;
        for (int l = 0; l < G; l++) s[l] = (G == 32) ? march_window<32>(p, t, l, &nxt, &wi) : march_window<8>(p, t, l, &nxt, &wi);
        if (wi.closed) {
            for (int v = 0; v < G; v++) {
                for (int k = 0; k <= 80; k++) {
                    float tt = k % 2 == 0 ? (v + k / 2 < G ? s[v + k / 2] : s[G - 1] + p.dt0 * (float)(v + k / 2 - G + 1))
                                          : s[v] + p.dt0 * (0.5f * (float)k + 0.01f * (float)(k % 7));
                    if (tt < s[v]) tt = s[v];
                    int want = G;
                    for (int l = v + 1; l < G; l++)
                        if (s[l] >= tt) { want = l; break; }
                    if (march_jump(wi, v, tt, G) != want) bad++;
                }
                if (march_jump(wi, v, INFINITY, G) != G) bad++;
                if (march_jump(wi, v, s[v] * 4.0f, G) != G) bad++;  // a later binade
            }
        }
        t = nxt;
    }
    return bad;
}

// march_fast_forward against the definition: the returned t must be the k-th sequential sum for some k >= 0, every skipped member
// must lie strictly below pend, and either the returned member is the first one >= pend or it is the last member of its binade /
// the function declined (returned t).  Returns the number of violations over a sweep of skip distances from t.
uint64_t mch_check_fast_forward(float t, uint32_t max_steps, uint32_t C, uint32_t H, int G) {
    const MarchParams p = make_march_params(16.0f, 0.0f, max_steps, C, H);
    uint64_t bad = 0;
    for (int d = 0; d <= 400; d += 3) {
        const float pend = t + p.dt0 * ((float)d + 0.37f);
        const float r = march_fast_forward(p, t, pend, G);
        if (r == t) continue;  // declined: always valid
        float s = t;
        int k = 0;
        bool found = false;
        for (; k < 100000; k++) {
            if (s == r) { found = true; break; }
            if (s >= pend) break;  // walked past the target without meeting the returned value
            s = s + p.dt0;
        }
        if (!found) { bad++; continue; }
        // s == r is member k; all members before it were < pend by the loop; r itself is >= pend unless the binade ended
        if (r < pend) {
            const float next = r + p.dt0;
            union { float f; uint32_t u; } a, b;
            a.f = r; b.f = next;
            if ((a.u >> 23) == (b.u >> 23)) {
                // still inside the binade: then the function must have stopped because the closed form caps at the last member whose
                // successor leaves the binade lattice; accept only if fewer than one window remains to the boundary
                union { float f; uint32_t u; } e; e.u = ((a.u >> 23) + 1u) << 23;
                if ((e.f - r) > p.dt0 * (float)(G + 1)) bad++;
            }
        }
    }
    return bad;
}

// Emulated group march of N rays (G > 0: serial resolve, G < 0: jump-table resolve with |G| lanes); counts[n] and, concatenated in ray order, the visited t values (ts, capacity cap).
// Returns the total number of samples.
uint64_t mch_group_march(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars,
                         const float* noises, int G, uint32_t* counts, float* ts, uint64_t cap) {
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H);
    std::vector<float> tl(max_steps);
    uint64_t total = 0;
    const uint32_t max_emit = g_max_emit ? g_max_emit : max_steps;
    for (uint32_t n = 0; n < N; n++) {
        const Ray r = make_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        const float near = nears[n];
        const float t0 = f_fma(f_clamp(f_mul(near, p.dt_gamma), p.dt_min, p.dt_max), noises[n], near);
        uint32_t c;
        if (G == 32) c = emulate_group<32>(p, r, grid, t0, fars[n], max_emit, tl.data());
        else if (G == 8) c = emulate_group<8>(p, r, grid, t0, fars[n], max_emit, tl.data());
        else if (G == -32) c = emulate_group_jump<32>(p, r, grid, t0, fars[n], max_emit, tl.data(), &g_jump_windows);
        else if (G == -8) c = emulate_group_jump<8>(p, r, grid, t0, fars[n], max_emit, tl.data(), &g_jump_windows);
        else c = emulate_group_jump<4>(p, r, grid, t0, fars[n], max_emit, tl.data(), &g_jump_windows);
        counts[n] = c;
        for (uint32_t k = 0; k < c && total + k < cap; k++) ts[total + k] = tl[k];
        total += c;
    }
    return total;
}
}
