"""CPU tier: the round-schedule theory of DESIGN.md section 3, checked on the ORACLE (the CPU restatement of the reference's
marcher and compositor, raymarching.cu:722-1035) -- no CUDA involved.

The reference's inference loop (renderer.py:353-379) takes n_step = clamp(N // n_alive, 1, 8) samples per alive ray per round.
The product renders on a faster schedule and claims the reference's bits.  What that rests on, each checked here with the
reference's own kernels:

  1. per-ray independence: a ray's result depends on the n_step SEQUENCE only, not on which other rays are rendered with it --
     any subset of rays on the reference's sequence reproduces the full-frame result bit for bit;
  2. the sequence is a function of the rays' death samples: NeRFNetwork._reference_sequence(histogram) == the sequence the loop took;
  3. rays whose result differs between two schedules exist on a scene with cameras inside the volume (so the problem is real), while
     every ray that comes out the same on both schedules also dies at the same sample (its death sample is schedule-independent);
     and the product's exactness criterion -- "no delta emitted after round 0 was an inexact difference" -- is SOUND: every ray that
     differs carries the flag (computed here from the reference marcher's own t and last_t) or was cut off by the max_steps cap,
     the one other way a schedule leaves a trace (the product detects that too and falls back);
  4. the fix-up: start from the fast schedule's death samples, re-render the differing rays on the sequence reconstructed from the
     histogram, iterate to a fixed point -> the reference schedule's results, bit for bit, for every ray.
"""
import numpy as np
import pytest

from cases import scene, scene_rays

pyo = pytest.importorskip("oracle.pyoracle")


def _field(xyzs):
    """A deterministic 'network': density and colour as smooth functions of the position (schedule-independent by construction)."""
    x, y, z = xyzs[:, 0].astype(np.float32), xyzs[:, 1].astype(np.float32), xyzs[:, 2].astype(np.float32)
    sig = (6.0 * (0.5 + 0.5 * np.sin(np.float32(3.1) * x + np.float32(1.7) * y - np.float32(2.3) * z))).astype(np.float32)
    rgb = np.stack([0.5 + 0.5 * np.sin(2 * x), 0.5 + 0.5 * np.cos(3 * y), 0.5 + 0.5 * np.sin(x + z)], -1).astype(np.float32)
    return sig, rgb


class _Frame:
    def __init__(self, name, n, seed):
        self.sc, self.ro, self.rd, _ = scene_rays(name, n, seed)
        self.nears, self.fars = pyo.near_far_from_aabb(self.ro, self.rd, self.sc.aabb, self.sc.min_near)
        self.n = n

    def render(self, schedule, subset=None, max_steps=1024, T=1e-4):
        """schedule: "reference" | "fast" | list of n_step (prescribed).  Returns dict(ws, depth, image, steps, seq)."""
        sc = self.sc
        idx = np.arange(self.n, dtype=np.int32) if subset is None else np.asarray(subset, np.int32)
        ro, rd, nears, fars = self.ro[idx], self.rd[idx], self.nears[idx], self.fars[idx]
        N = len(idx)
        alive = np.arange(N, dtype=np.int32)
        rays_t = nears.copy()
        ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
        steps = np.zeros(N, np.int64)
        flags = np.zeros(N, bool)   # the product's criterion (csrc/raymarch.cu kCtlInexact): an inexact delta after round 0
        seq, step, r = [], 0, 0
        while step < max_steps and len(alive) > 0:
            if isinstance(schedule, list):
                if r >= len(schedule):
                    break
                n_step = schedule[r]
            elif schedule == "reference":
                n_step = max(min(N // len(alive), 8), 1)          # renderer.py:357
            else:
                n_step = 1 if r == 0 else 32                       # "fast": far fewer, far longer rounds
            M_rows = len(alive) * n_step
            M_rows += 128 - M_rows % 128
            xyzs, dirs, deltas, inexact = pyo.march_rays_track(len(alive), n_step, alive, rays_t, ro, rd, sc.bound, sc.density_bitfield, sc.cascade,
                                                               128, nears, fars, np.zeros(len(alive), np.float32), M_rows, 0.0, max_steps)
            if r > 0:
                flags[alive] |= inexact.astype(bool)
            sig, rgb = _field(xyzs)
            alive2, rays_t, ws, depth, image, done = pyo.composite_rays_steps(len(alive), n_step, alive, rays_t, sig, rgb, deltas, ws, depth, image, T)
            steps[alive] += done
            alive = alive2[alive2 >= 0]                            # renderer.py:375
            seq.append(n_step)
            step += n_step
            r += 1
        cut = np.zeros(N, bool)
        cut[alive] = True   # rays the max_steps cap cut off (still alive when the loop ended)
        return dict(ws=ws, depth=depth, image=image, steps=steps, seq=seq, flags=flags, cut=cut)


@pytest.mark.parametrize("name,n", [("bonsai", 700), ("lego", 500)])
def test_round_schedule_theory_on_the_oracle(name, n):
    from laenerf_b200.nerf import NeRFNetwork
    fr = _Frame(name, n, 11)
    ref = fr.render("reference")
    fast = fr.render("fast")
    assert len(fast["seq"]) < len(ref["seq"]) / 2
    same = (ref["ws"] == fast["ws"]) & (ref["depth"] == fast["depth"]) & (ref["image"] == fast["image"]).all(1)
    # 3. schedule-independent rays die at the same sample; on the bonsai shape some rays DO depend on the schedule
    assert (ref["steps"][same] == fast["steps"][same]).all()
    if name == "bonsai":
        assert 0 < (~same).sum() < n // 4, int((~same).sum())
    # 3b. the product's criterion is sound: every ray that differs was flagged by the FAST pass (an inexact delta after round 0);
    # rays without a flag come out the same on both schedules.  (The flag may over-report: it is a sufficient condition.)
    # A ray that is still alive when the loop runs into max_steps is cut at a schedule-dependent sample (the rounds hand out 1024-1031
    # samples on one schedule, 1025 on the other): the product sends such a frame to the reference schedule (ctl[14], "cap cut").
    cut = ref["cut"] | fast["cut"]
    assert not (~same & ~fast["flags"] & ~cut).any(), int((~same & ~fast["flags"] & ~cut).sum())
    if name == "lego":
        assert not fast["flags"].any() and same.all()      # cameras outside the volume: the frame is provably schedule-independent
    # 2. the sequence follows from the histogram of the death samples
    cap = 1024 + 72
    hist = np.bincount(np.minimum(ref["steps"], cap), minlength=cap + 1)
    assert NeRFNetwork._reference_sequence(hist, n, 1024) == ref["seq"]
    # 1. any subset on the reference's sequence reproduces the full-frame result
    rng = np.random.default_rng(0)
    subset = np.unique(np.concatenate([np.nonzero(~same)[0], rng.choice(n, 40, replace=False)]))
    part = fr.render(list(ref["seq"]), subset=subset)
    for k in ("ws", "depth", "image", "steps"):
        assert np.array_equal(part[k], ref[k][subset]), k
    # 4. the fix-up from the FAST pass's information alone
    bad = np.nonzero(~same)[0]
    steps = fast["steps"].copy()
    out = {k: fast[k].copy() for k in ("ws", "depth", "image")}
    for _ in range(4):
        hist = np.bincount(np.minimum(steps, cap), minlength=cap + 1)
        seq = NeRFNetwork._reference_sequence(hist, n, 1024)
        if len(bad) == 0:
            break
        redo = fr.render(seq, subset=bad)
        steps[bad] = redo["steps"]
        for k in out:
            out[k][bad] = redo[k]
        hist2 = np.bincount(np.minimum(steps, cap), minlength=cap + 1)
        if NeRFNetwork._reference_sequence(hist2, n, 1024) == seq:
            break
    else:
        pytest.fail("no fixed point")
    assert seq == ref["seq"]
    for k in out:
        assert np.array_equal(out[k], ref[k]), k


def _occupied_box(sc):
    """numpy mirror of csrc/occupancy.cu k_occupied_box: box around every occupied cell of every cascade, two cells of margin."""
    H = 128
    lo_w, hi_w = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
    bits = np.unpackbits(sc.density_bitfield, bitorder="little").reshape(sc.cascade, H ** 3)
    coords = pyo.morton3D_invert(np.arange(H ** 3, dtype=np.int32))
    for c in range(sc.cascade):
        occ = coords[bits[c] == 1]
        if len(occ) == 0:
            continue
        half = np.float32(min(2 ** c, sc.bound))
        cell = np.float32(2.0) * half / np.float32(H)
        lo_w = np.minimum(lo_w, -half + (occ.min(0).astype(np.float32) - np.float32(2.0)) * cell)
        hi_w = np.maximum(hi_w, -half + (occ.max(0).astype(np.float32) + np.float32(3.0)) * cell)
    return np.concatenate([lo_w, hi_w]).astype(np.float32)


def _clip_far(box, ro, rd, fars):
    """numpy mirror of csrc/march_core.cuh clip_far_to_box (float32 slab test; a miss gives -inf)."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        rdi = (np.float32(1.0) / rd.astype(np.float32)).astype(np.float32)
        t1 = ((box[None, :3] - ro) * rdi).astype(np.float32)
        t2 = ((box[None, 3:] - ro) * rdi).astype(np.float32)
        t_in = np.fmax.reduce(np.fmin(t1, t2), axis=1)     # fmin / fmax drop NaN, as the device's fminf / fmaxf do
        t_out = np.fmin.reduce(np.fmax(t1, t2), axis=1)
    out = np.fmin(fars, t_out).astype(np.float32)
    out[~(t_in <= t_out)] = -np.inf
    return out


@pytest.mark.parametrize("name", ["lego", "flower", "bonsai"])
def test_clipping_far_to_the_occupied_box_is_exact_on_the_oracle(name):
    """DESIGN.md section 3, economy (2): the reference's marchers with `far` clipped to the exit of the box around the occupied cells
    emit the same samples -- training march (counts, offsets, positions, deltas) and the inference loop (image, depth, weights, death
    samples) -- and a good part of the rays of the object-centred scene end much earlier."""
    n = 400
    fr = _Frame(name, n, 23)
    sc = fr.sc
    box = _occupied_box(sc)
    clipped = _clip_far(box, fr.ro, fr.rd, fr.fars)
    assert (clipped <= fr.fars).all()
    if name == "lego":
        assert (clipped < fr.fars - 0.2).mean() > 0.5        # most rays leave the object's box long before the scene box
    rng = np.random.default_rng(3)
    noises = rng.random(n, dtype=np.float32)
    M = n * 1024
    for dt_gamma in (0.0, 1.0 / 256):
        a = pyo.march_rays_train(fr.ro, fr.rd, sc.density_bitfield, sc.bound, dt_gamma, 1024, sc.cascade, 128, M, fr.nears, fr.fars, noises)
        b = pyo.march_rays_train(fr.ro, fr.rd, sc.density_bitfield, sc.bound, dt_gamma, 1024, sc.cascade, 128, M, fr.nears, clipped, noises)
        total = int(a[4][0])
        assert total == int(b[4][0]) and total > 100
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x[:total], y[:total])
        assert np.array_equal(a[3], b[3])
    ref = fr.render("reference")
    keep = fr.fars
    fr.fars = clipped
    try:
        cl = fr.render("reference")
    finally:
        fr.fars = keep
    assert cl["seq"] == ref["seq"]
    for k in ("ws", "depth", "image", "steps"):
        assert np.array_equal(cl[k], ref[k]), k


def test_one_pass_rerender_equals_the_round_by_round_loop_on_the_oracle():
    """What lnrf_march_rays_prescribed + lnrf_composite_rays_prescribed do, restated with the oracle's kernels: a
    ray is marched through ALL rounds of a given n_step sequence on its own -- each round starting from the t the compositor would have
    rebuilt, rays_t + the round's deltas[.][1] added one by one (raymarching.cu:1006) -- the 'network' runs once over those samples,
    and the compositor walks them in ONE call.  Image, depth, weights and death sample equal the round-by-round loop's, bit for bit."""
    fr = _Frame("bonsai", 300, 5)
    sc = fr.sc
    ref = fr.render("reference")
    seq = ref["seq"]
    rng = np.random.default_rng(1)
    for i in rng.choice(fr.n, 60, replace=False):
        ro, rd = fr.ro[i:i + 1], fr.rd[i:i + 1]
        nears, fars = fr.nears[i:i + 1], fr.fars[i:i + 1]
        alive = np.zeros(1, np.int32)
        t = nears.copy()
        xs, ds = [], []
        for n_step in seq:                                   # the marcher alone, round after round
            M_rows = n_step + 128 - n_step % 128
            xyzs, _, deltas = pyo.march_rays(1, n_step, alive, t, ro, rd, sc.bound, sc.density_bitfield, sc.cascade, 128, nears, fars,
                                             np.zeros(1, np.float32), M_rows, 0.0, 1024)
            cnt = int((deltas[:n_step, 0] != 0).sum())
            xs.append(xyzs[:cnt]); ds.append(deltas[:cnt])
            for k in range(cnt):                             # what composite_rays leaves in rays_t
                t = (t + deltas[k, 1]).astype(np.float32)
            if cnt < n_step:
                break
        xyz = np.concatenate(xs) if xs else np.zeros((0, 3), np.float32)
        dl = np.concatenate(ds) if ds else np.zeros((0, 2), np.float32)
        total = len(xyz)
        if total == 0:
            assert ref["ws"][i] == 0 and ref["steps"][i] == 0
            continue
        sig, rgb = _field(xyz)
        pad = np.zeros((1, 2), np.float32)                   # one empty slot behind the samples: the loop's `deltas == 0` exit
        _, _, ws, depth, image, done = pyo.composite_rays_steps(1, total + 1, alive, nears.copy(), np.append(sig, np.float32(0)),
                                                                np.concatenate([rgb, np.zeros((1, 3), np.float32)]),
                                                                np.concatenate([dl, pad]), np.zeros(1, np.float32), np.zeros(1, np.float32),
                                                                np.zeros((1, 3), np.float32), 1e-4)
        assert ws[0] == ref["ws"][i] and depth[0] == ref["depth"][i] and np.array_equal(image[0], ref["image"][i]), i
        assert int(done[0]) == int(ref["steps"][i]), i
