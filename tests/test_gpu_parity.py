"""GPU tier: liblaenerf_b200.so (called through the C ABI) against (a) the golden vectors frozen from the
reference's own CUDA extensions and (b) the CPU oracle, on the seeded cases of tests/cases.py.
Bit-exact for counts / offsets / positions / bitfields / indices; stated tolerances for floating point (cases.TOL)."""
import numpy as np
import pytest

from backends import canonical_rays
from cases import CASES, compare, run_case, scene_rays
from conftest import golden_scales, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_matches_reference_golden(name, ours_backend):
    gold = load_golden(name)
    if gold is None:
        pytest.skip(f"tests/golden/ref_{name}.npz not generated yet")
    got = run_case(name, ours_backend)
    compare(got, gold)


@pytest.mark.parametrize("name", list(CASES))
def test_matches_cpu_oracle(name, ours_backend):
    from backends import OracleBackend
    got = run_case(name, ours_backend)
    scales = {}
    if name == "grid_small":
        from cases import grid_config
        scales[(8, 16)] = ours_backend.grid_level_scales(8, grid_config(8, 2, 3, 16, 12, 512)[1], 16)
    if name == "grid_d2c4":
        from cases import grid_config
        scales[(4, 16)] = ours_backend.grid_level_scales(4, grid_config(4, 4, 2, 16, 10, 128)[1], 16)
    want = run_case(name, OracleBackend(device_scales=scales))
    from cases import TOL_VS_ORACLE
    compare(got, want, table=TOL_VS_ORACLE)


def test_march_overflow_drops_rays_like_the_reference(ours_backend, oracle_backend):
    """M smaller than the total: rays with offset + count > M write nothing and the tail is zero (raymarching.cu:416)."""
    sc, ro, rd, rng = scene_rays("lego", 512, 31)
    nears, fars = oracle_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    noises = rng.random(512, dtype=np.float32)
    full = oracle_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, 0.0, 1024, 1, 128, 512 * 1024, nears, fars, noises)
    M = int(full[4][0]) // 2 + 5
    for counter0 in (0, 40):
        cnt = np.array([counter0, 3], np.int32)
        want = oracle_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, 0.0, 1024, 1, 128, M, nears, fars, noises, cnt)
        got = ours_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, 0.0, 1024, 1, 128, M, nears, fars, noises, cnt)
        for a, b, nm in zip(got, want, ("xyzs", "dirs", "deltas", "rays", "counter")):
            assert np.array_equal(a, b), nm


def test_march_empty_and_tiny_inputs(ours_backend, oracle_backend):
    sc, ro, rd, rng = scene_rays("lego", 3, 32)
    nears, fars = ours_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    empty = np.zeros_like(sc.density_bitfield)
    x, d, dl, rays, cnt = ours_backend.march_train(ro, rd, empty, sc.bound, 0.0, 1024, 1, 128, 256, nears, fars, np.zeros(3, np.float32))
    assert cnt[0] == 0 and cnt[1] == 3 and not x.any() and not dl.any() and np.array_equal(rays[:, 2], [0, 0, 0])
    full = np.full_like(sc.density_bitfield, 255)  # every cell occupied: rays emit until far or the max_steps budget
    for max_steps in (16, 64, 1024):
        got = ours_backend.march_train(ro, rd, full, sc.bound, 0.0, max_steps, 1, 128, 3 * max_steps, nears, fars, np.zeros(3, np.float32))
        want = oracle_backend.march_train(ro, rd, full, sc.bound, 0.0, max_steps, 1, 128, 3 * max_steps, nears, fars, np.zeros(3, np.float32))
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    assert (got[3][:, 2] > 100).all()


def test_composite_backward_zero_fill_equals_reference_contract(ours_backend):
    sc, ro, rd, rng = scene_rays("lego", 300, 33)
    nears, fars = ours_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    x, d, dl, rays, cnt = ours_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, 0.0, 1024, 1, 128, 300 * 1024, nears, fars,
                                                   rng.random(300, dtype=np.float32))
    total = int(cnt[0])
    M = total + (128 - total % 128)
    sig = (rng.random(M, dtype=np.float32) * 300).astype(np.float32)
    rgb = rng.random((M, 3), dtype=np.float32)
    ws, depth, image = ours_backend.composite_train_fwd(sig, rgb, dl[:M], rays, 1e-4)
    gws, gimg = rng.standard_normal(300).astype(np.float32), rng.standard_normal((300, 3)).astype(np.float32)
    a = ours_backend.composite_train_bwd(gws, gimg, sig, rgb, dl[:M], rays, ws, image, 1e-4, zero_fill=False)
    b = ours_backend.composite_train_bwd(gws, gimg, sig, rgb, dl[:M], rays, ws, image, 1e-4, zero_fill=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (a[0] == 0).sum() > M - total  # early-terminated tails are zero as well


def test_compact_alive(ours_backend):
    rng = np.random.default_rng(34)
    for n in (1, 31, 1024, 1025, 5000, 640000):
        ra = rng.integers(-1, 1 << 20, size=n).astype(np.int32)
        ra[rng.random(n) < 0.4] = -1
        out, k = ours_backend.compact(ra)
        assert k == int((ra >= 0).sum()) and np.array_equal(out, ra[ra >= 0])
    out, k = ours_backend.compact(np.full(777, -1, np.int32))
    assert k == 0


def test_grid_layouts_agree(ours_backend):
    from cases import grid_config
    rng = np.random.default_rng(35)
    offsets, pls = grid_config(8, 2, 3, 16, 12, 512)
    x = rng.random((1000, 3), dtype=np.float32)
    emb = rng.uniform(-1, 1, size=(int(offsets[-1]), 2)).astype(np.float32)
    blc = ours_backend.grid_fwd(x, emb, offsets, pls, 16)                 # tiled hot kernel
    lbc = ours_backend.grid_fwd(x, emb, offsets, pls, 16, layout=0)       # generic kernel, reference-native layout
    assert np.allclose(blc, lbc.transpose(1, 0, 2).reshape(1000, 16), rtol=1e-6, atol=1e-7)
    g = rng.standard_normal((1000, 16)).astype(np.float32)
    a = ours_backend.grid_bwd(g, x, offsets, 2, pls, 16)
    b = ours_backend.grid_bwd(np.ascontiguousarray(g.reshape(1000, 8, 2).transpose(1, 0, 2)), x, offsets, 2, pls, 16, layout=0)
    assert np.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_ffmlp_inference_equals_training_forward(ours_backend):
    rng = np.random.default_rng(36)
    for in_dim, nl in ((32, 2), (32, 3), (64, 2), (16, 4), (48, 2)):
        w = (rng.uniform(-0.2, 0.2, 64 * (in_dim + 64 * (nl - 1) + 16))).astype(np.float16).astype(np.float32)
        x = (rng.standard_normal((640, in_dim)) * 0.5).astype(np.float16).astype(np.float32)
        out, fb = ours_backend.ffmlp_fwd(x, w, in_dim, 16, 64, nl)
        inf = ours_backend.ffmlp_fwd(x, w, in_dim, 16, 64, nl, inference=True)
        assert np.array_equal(out, inf)
        # plain fp32 reference of the same op with fp16 storage of the activations
        h = x
        W0 = w[:64 * in_dim].reshape(64, in_dim)
        h = np.maximum(h @ W0.T, 0).astype(np.float16).astype(np.float32)
        assert np.allclose(fb[0], h, rtol=2e-3, atol=2e-3)
        off = 64 * in_dim
        for l in range(1, nl):
            Wl = w[off:off + 4096].reshape(64, 64)
            off += 4096
            h = np.maximum(h @ Wl.T, 0).astype(np.float16).astype(np.float32)
            assert np.allclose(fb[l], h, rtol=4e-3, atol=4e-3)
        y = h @ w[off:].reshape(16, 64).T
        assert np.allclose(out, y, rtol=1e-2, atol=1e-2)


class _env:
    def __init__(self, key, value):
        self.key, self.value = key, str(value)

    def __enter__(self):
        import os
        self.old = os.environ.get(self.key)
        os.environ[self.key] = self.value

    def __exit__(self, *a):
        import os
        if self.old is None:
            os.environ.pop(self.key, None)
        else:
            os.environ[self.key] = self.old


@pytest.mark.parametrize("name", ["march_lego", "march_flower", "march_bonsai", "infer_lego", "distill_flower"])
def test_jump_table_resolve_is_bit_identical_to_the_serial_resolve(name, ours_backend):
    """LNRF_MARCH_JUMP=0 (serial ballot loop per window) and the default jump-table resolve (march_jump + pointer doubling)
    must produce the same bits for every marcher (training 32 lanes, inference / distill 4 and 8 lanes)."""
    if name not in CASES:
        pytest.skip(f"no case {name}")
    with _env("LNRF_MARCH_JUMP", 0):
        a = run_case(name, ours_backend)
    with _env("LNRF_MARCH_JUMP", 1):
        b = run_case(name, ours_backend)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_march_budget_and_far_edges_serial_vs_jump(ours_backend, oracle_backend):
    """Rays that end on the sample budget inside a window (tiny max_steps) and at `far`, against the sequential oracle, with the
    jump-table resolve on and off."""
    sc, ro, rd, rng = scene_rays("lego", 700, 37)
    nears, fars = oracle_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    noises = rng.random(700, dtype=np.float32)
    full = np.full_like(sc.density_bitfield, 255)
    for grid in (sc.density_bitfield, full):
        for max_steps in (7, 48, 333, 1024):
            want = oracle_backend.march_train(ro, rd, grid, sc.bound, 0.0, max_steps, 1, 128, 700 * max_steps, nears, fars, noises)
            for jump in (0, 1):
                with _env("LNRF_MARCH_JUMP", jump):
                    got = ours_backend.march_train(ro, rd, grid, sc.bound, 0.0, max_steps, 1, 128, 700 * max_steps, nears, fars, noises)
                for a, b, nm in zip(got, want, ("xyzs", "dirs", "deltas", "rays", "counter")):
                    assert np.array_equal(a, b), (nm, max_steps, jump)


def test_paired_table_access_equals_unpaired(ours_backend):
    """LNRF_GRID_PAIR: x / x+1 corners in one 8-byte gather / one REDG.F16x4 (default) against one access per corner.  The
    forward is the same arithmetic on the same values (bit-equal); the backward differs by the order of fp16 atomic adds only."""
    from cases import grid_config
    rng = np.random.default_rng(38)
    offsets, pls = grid_config(16, 2, 3, 16, 19, 2048)
    x = rng.random((4096 + 77, 3), dtype=np.float32)
    x[:1000] = np.linspace(0.2, 0.6, 1000, dtype=np.float32)[:, None] * np.array([1.0, 0.7, 0.3], np.float32) + 0.1  # ray-like runs
    emb = rng.uniform(-1, 1, size=(int(offsets[-1]), 2)).astype(np.float16).astype(np.float32)
    g = (rng.standard_normal((x.shape[0], 32)) * 0.05).astype(np.float16).astype(np.float32)
    for half in (True, False):
        with _env("LNRF_GRID_PAIR", 0):
            f0 = ours_backend.grid_fwd(x, emb, offsets, pls, 16, half=half)
            b0 = ours_backend.grid_bwd(g, x, offsets, 2, pls, 16, half=half)
        with _env("LNRF_GRID_PAIR", 3):
            f1 = ours_backend.grid_fwd(x, emb, offsets, pls, 16, half=half)
            b1 = ours_backend.grid_bwd(g, x, offsets, 2, pls, 16, half=half)
        assert np.array_equal(f0, f1), half
        assert np.abs(b0).max() > 0
        assert np.allclose(b0, b1, rtol=2e-2, atol=2e-3) if half else np.allclose(b0, b1, rtol=1e-4, atol=1e-5), half


def test_sph_from_ray_matches_oracle():
    """raymarching.sph_from_ray (raymarching.cu:162-209; no LAENeRF config calls it, the binding exists) against the C oracle:
    atan2f / sqrtf on the device are not bit-identical to libm, measured 2.1e-7 on the lego-shape rays."""
    import torch
    from oracle import pyoracle
    from laenerf_b200 import raymarching
    sc, ro, rd, rng = scene_rays("lego", 3000, 51)
    want = pyoracle.sph_from_ray(ro, rd, 4.0)
    got = raymarching.sph_from_ray(torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda(), 4.0).cpu().numpy()
    assert got.shape == want.shape == (3000, 2) and np.isfinite(got).all()
    assert float(np.abs(got - want).max()) <= 2e-6
