"""CPU tier: the marching core shared by the CUDA kernels (laenerf_b200/csrc/march_core.cuh), compiled for the host.

1. the closed-form window generator (one fma per lane) against plain sequential float adds -- the definition of the
   reference's t-sequence (raymarching.cu:396-398, 455);
2. the lane-group resolve algorithm of raymarch.cu (emulated lane by lane in march_core_host.cpp) against the
   sequential oracle: identical per-ray sample counts and visited t values for 8- and 32-lane groups."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import scene_rays

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "laenerf_b200", "csrc")


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-s", "-C", CSRC, "_build/march_core_host.so"], check=True)
    lib = C.CDLL(os.path.join(CSRC, "_build", "march_core_host.so"))
    lib.mch_check_window.restype = C.c_uint64
    lib.mch_check_window.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
    lib.mch_group_march.restype = C.c_uint64
    lib.mch_group_march.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + \
        [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    return lib


@pytest.mark.parametrize("G", [8, 32])
@pytest.mark.parametrize("dt_gamma,max_steps,Cc", [(0.0, 1024, 1), (0.0, 1024, 5), (0.0, 333, 2), (1.0 / 256, 1024, 5), (0.0, 64, 1)])
def test_window_generator_is_the_sequential_sum(host, G, dt_gamma, max_steps, Cc):
    rng = np.random.default_rng(G * 1000 + max_steps)
    starts = np.concatenate([rng.uniform(0.05, 40.0, 300), [0.2, 0.5, 1.0, 2.0, 4.0, 7.99999, 1e-3, 0.0, 3.0517578125e-05],
                             np.exp2(rng.integers(-3, 5, 20)) - 1e-7]).astype(np.float32)
    for t0 in starts:
        assert host.mch_check_window(float(t0), dt_gamma, max_steps, Cc, 128, G, 40) == 0, f"t0={t0!r}"


@pytest.mark.parametrize("name,n,dt_gamma", [("lego", 384, 0.0), ("flower", 192, 0.0), ("flower", 128, 1.0 / 128)])
@pytest.mark.parametrize("G", [8, 32])
def test_group_resolve_equals_sequential_march(host, oracle_backend, name, n, dt_gamma, G):
    sc, ro, rd, rng = scene_rays(name, n, 21)
    nears, fars = oracle_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    noises = rng.random(n, dtype=np.float32)
    M = n * sc.max_steps
    xyzs, dirs, deltas, rays, counter = oracle_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, dt_gamma, sc.max_steps,
                                                                   sc.cascade, 128, M, nears, fars, noises)
    counts = np.zeros(n, np.uint32)
    ts = np.zeros(int(counter[0]) + 16, np.float32)
    grid = np.ascontiguousarray(sc.density_bitfield)
    total = host.mch_group_march(ro.ctypes.data, rd.ctypes.data, grid.ctypes.data, sc.bound, dt_gamma, sc.max_steps, n, sc.cascade, 128,
                                 nears.ctypes.data, fars.ctypes.data, noises.ctypes.data, G, counts.ctypes.data, ts.ctypes.data,
                                 ts.shape[0])
    assert total == int(counter[0])
    assert np.array_equal(counts.astype(np.int32), rays[:, 2])
    # visited t values reproduce the oracle's sample positions: x = clamp(o + t d)
    t = ts[:total]
    ray_of = np.repeat(np.arange(n), counts)
    x = np.clip((ro[ray_of].astype(np.float64) + t[:, None].astype(np.float64) * rd[ray_of].astype(np.float64)), -sc.bound, sc.bound)
    assert np.allclose(x, xyzs[:total], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name,n,dt_gamma,max_steps", [("lego", 384, 0.0, None), ("flower", 192, 0.0, None), ("bonsai", 192, 0.0, None),
                                                        ("flower", 96, 1.0 / 128, None), ("lego", 256, 0.0, 48), ("lego", 256, 0.0, 7),
                                                        ("bonsai", 128, 0.0, 12)])
@pytest.mark.parametrize("G", [4, 8, 32])
def test_jump_table_resolve_equals_sequential_march(host, oracle_backend, name, n, dt_gamma, max_steps, G):
    """The pointer-doubling resolve of closed-form windows (march_jump + orbit doubling, raymarch.cu march_group) against the
    sequential oracle: identical counts and visited t values, including rays that end on the sample budget (small max_steps)
    and at `far`; windows that are not closed-form (binade crossings, dt_gamma > 0) take the serial resolve."""
    host.mch_jump_windows.restype = C.c_uint64
    sc, ro, rd, rng = scene_rays(name, n, 33)
    ms = sc.max_steps if max_steps is None else max_steps
    nears, fars = oracle_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    noises = rng.random(n, dtype=np.float32)
    M = n * ms
    xyzs, dirs, deltas, rays, counter = oracle_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, dt_gamma, ms, sc.cascade, 128,
                                                                   M, nears, fars, noises)
    counts = np.zeros(n, np.uint32)
    ts = np.zeros(int(counter[0]) + 16, np.float32)
    grid = np.ascontiguousarray(sc.density_bitfield)
    before = host.mch_jump_windows()
    total = host.mch_group_march(ro.ctypes.data, rd.ctypes.data, grid.ctypes.data, sc.bound, dt_gamma, ms, n, sc.cascade, 128,
                                 nears.ctypes.data, fars.ctypes.data, noises.ctypes.data, -G, counts.ctypes.data, ts.ctypes.data,
                                 ts.shape[0])
    assert total == int(counter[0])
    assert np.array_equal(counts.astype(np.int32), rays[:, 2])
    if dt_gamma == 0.0:
        assert host.mch_jump_windows() > before  # the new path really ran
    if max_steps is not None:
        assert int(counts.max()) == ms  # some ray did end on the budget
    t = ts[:total]
    ray_of = np.repeat(np.arange(n), counts)
    x = np.clip((ro[ray_of].astype(np.float64) + t[:, None].astype(np.float64) * rd[ray_of].astype(np.float64)), -sc.bound, sc.bound)
    assert np.allclose(x, xyzs[:total], rtol=0, atol=1e-6 * max(1.0, sc.bound))


@pytest.mark.parametrize("G", [4, 8, 32])
def test_fast_forward_over_pending_skips_changes_nothing_but_the_window_count(host, oracle_backend, G):
    """march_fast_forward: on the 5-cascade bonsai shape an empty outer-cascade voxel spans ~37 members, so a pending skip used to
    cost several idle windows; with the fast-forward the emulated march needs far fewer windows and still reproduces the
    sequential oracle (counts and visited t) exactly."""
    host.mch_jump_windows.restype = C.c_uint64
    sc, ro, rd, rng = scene_rays("bonsai", 160, 35)
    nears, fars = oracle_backend.near_far(ro, rd, sc.aabb, sc.min_near)
    noises = rng.random(160, dtype=np.float32)
    want = oracle_backend.march_train(ro, rd, sc.density_bitfield, sc.bound, 0.0, sc.max_steps, sc.cascade, 128, 160 * sc.max_steps,
                                      nears, fars, noises)
    grid = np.ascontiguousarray(sc.density_bitfield)
    windows, results = [], []
    for on in (0, 1):
        host.mch_set_fast_forward(on)
        counts = np.zeros(160, np.uint32)
        ts = np.zeros(int(want[4][0]) + 16, np.float32)
        before = host.mch_jump_windows()
        total = host.mch_group_march(ro.ctypes.data, rd.ctypes.data, grid.ctypes.data, sc.bound, 0.0, sc.max_steps, 160, sc.cascade, 128,
                                     nears.ctypes.data, fars.ctypes.data, noises.ctypes.data, -G, counts.ctypes.data, ts.ctypes.data,
                                     ts.shape[0])
        windows.append(host.mch_jump_windows() - before)
        results.append((total, counts.copy(), ts[:total].copy()))
    host.mch_set_fast_forward(1)
    for total, counts, ts in results:
        assert total == int(want[4][0]) and np.array_equal(counts.astype(np.int32), want[3][:, 2])
    assert np.array_equal(results[0][2], results[1][2])
    assert windows[1] * (2 if G < 32 else 1) < windows[0], windows  # a 32-member window already spans most of a 37-member voxel


@pytest.mark.parametrize("G", [8, 32])
@pytest.mark.parametrize("max_steps,Cc", [(1024, 1), (1024, 5), (333, 2), (64, 1)])
def test_march_jump_equals_linear_search(host, G, max_steps, Cc):
    """march_jump (integer arithmetic on mantissa fields) against "first later member with s >= tt" by linear search, for every
    lane of closed-form windows and targets on members, between members, beyond the window, in a later binade and at infinity."""
    host.mch_check_jump.restype = C.c_uint64
    host.mch_check_jump.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
    rng = np.random.default_rng(G + max_steps)
    starts = np.concatenate([rng.uniform(0.05, 40.0, 100), [0.2, 0.5, 1.0, 2.0, 3.99, 4.0, 7.99999, 15.9]]).astype(np.float32)
    for t0 in starts:
        assert host.mch_check_jump(float(t0), max_steps, Cc, 128, G, 12) == 0, f"t0={t0!r}"


@pytest.mark.parametrize("G", [4, 8, 32])
def test_march_fast_forward_lands_on_the_sequence(host, G):
    """march_fast_forward returns a member of the sequential t-sequence, never skips a member >= the pending target, and stops
    at the target or at the end of the binade."""
    host.mch_check_fast_forward.restype = C.c_uint64
    host.mch_check_fast_forward.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    rng = np.random.default_rng(G)
    starts = np.concatenate([rng.uniform(0.05, 50.0, 150), [0.2, 1.0, 1.99, 3.9, 4.0, 7.9, 15.99, 31.5]]).astype(np.float32)
    for t0 in starts:
        assert host.mch_check_fast_forward(float(t0), 1024, 5, 128, G) == 0, f"t0={t0!r}"
