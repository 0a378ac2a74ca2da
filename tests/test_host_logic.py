"""CPU tier: host-side logic of the drop-in modules (no kernels run): table sizing, parameter layouts, padding rules,
alias packages, and the hot path's product code never importing the oracle."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grid_encoder_table_sizing_matches_reference_formula():
    from laenerf_b200.gridencoder import GridEncoder
    from oracle import pyoracle
    for bound, expect_head in ((1, [4920, 13824, 32768, 85184, 216000]), (2, [4920, 15632, 42880, 125000, 373248])):
        enc = GridEncoder(desired_resolution=2048 * bound)
        off = enc.offsets.numpy()
        sizes = np.diff(off)
        assert list(sizes[:5]) == expect_head
        assert all(s == 524288 for s in sizes[5:])
        ref_off, pls = pyoracle.grid_offsets(desired_resolution=2048 * bound)
        assert np.array_equal(off, ref_off) and abs(pls - enc.per_level_scale) < 1e-12
        assert enc.embeddings.shape == (off[-1], 2) and enc.embeddings.dtype == torch.float32
        assert enc.output_dim == 32 and float(enc.embeddings.abs().max()) <= 1e-4
    assert GridEncoder(desired_resolution=2048).offsets[-1].item() == 6119864


def test_ffmlp_parameter_layout_and_seed_side_effect():
    from laenerf_b200.ffmlp import FFMLP
    torch.manual_seed(7)
    a = FFMLP(32, 16, 64, 2)
    assert a.num_parameters == 64 * (32 + 64 + 16) == a.weights.numel()
    b = FFMLP(32, 3, 64, 3)
    assert b.padded_output_dim == 16 and b.num_parameters == 64 * (32 + 2 * 64 + 16)
    # construction reseeds the global RNG with 42 (ffmlp.py:142): the next draw is the same after either constructor
    FFMLP(32, 16, 64, 2)
    x = torch.rand(3)
    FFMLP(32, 16, 64, 2)
    assert torch.equal(x, torch.rand(3))
    assert float(a.weights.abs().max()) <= (3 / 64) ** 0.5
    with pytest.raises(AssertionError):
        FFMLP(30, 16, 64, 2)
    with pytest.raises(AssertionError):
        FFMLP(32, 17, 64, 2)
    with pytest.raises(AssertionError):
        FFMLP(32, 16, 64, 1)


def test_dropin_alias_packages_resolve():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        import raymarching
        from ffmlp import FFMLP
        from gridencoder import GridEncoder
        from raymarching import raymarching as rm_mod
        from shencoder import SHEncoder
        for fn in ("near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                   "composite_rays_train", "march_rays", "march_rays_distill", "composite_rays", "composite_rays_distill"):
            assert callable(getattr(raymarching, fn)) and callable(getattr(rm_mod, fn))
        assert GridEncoder.__name__ == "GridEncoder" and FFMLP.__name__ == "FFMLP" and SHEncoder.__name__ == "SHEncoder"
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "laenerf_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|pyoracle|liblaenerf_oracle|oracle/_ref|oracle\.c", text):
                    bad.append(os.path.join(d, f))
    for d in ("dropin",):
        for dd, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith(".py") and "oracle" in open(os.path.join(dd, f)).read():
                    bad.append(os.path.join(dd, f))
    assert not bad, bad


def test_scene_generator_shapes():
    from cases import scene
    sc = scene("lego")
    assert sc.cascade == 1 and sc.density_bitfield.shape == (128 ** 3 // 8,) and sc.density_grid.shape == (1, 128 ** 3)
    assert 0.005 < sc.occupancy_fraction() < 0.2
    from laenerf_b200.scene import packbits_np
    from oracle import pyoracle
    assert np.array_equal(packbits_np(sc.density_grid, 10.0), pyoracle.packbits(sc.density_grid, 10.0))


def test_style_encoder_module_topology_without_a_gpu():
    """Row a-13 host logic: the LAENeRF mirror sizes its two nets like tcnn does (inputs padded to 16, 3 matmuls for
    num_layers = 3) and refuses what is out of scope instead of silently training without it."""
    from types import SimpleNamespace
    from laenerf_b200.style_encoder import LAENeRF
    with pytest.raises(RuntimeError):
        LAENeRF(SimpleNamespace(bound=2, num_palette_bases=8, style_weight=1.0), device="cpu")
    with pytest.raises(RuntimeError):
        LAENeRF(SimpleNamespace(bound=2, num_palette_bases=17, style_weight=0.0), device="cpu")
    m = LAENeRF(SimpleNamespace(bound=2, num_palette_bases=8, style_weight=0.0), dir_encoding="sphere_harmonics", device="cpu")
    assert m.in_dim == 32 and m.in_dim_dir == 9 and m.offset_in_dim == 48
    assert (m.offset_net.input_dim, m.offset_net.num_layers, m.offset_net.output_dim) == (48, 2, 3)
    assert (m.weight_net.input_dim, m.weight_net.num_layers, m.weight_net.output_dim) == (32, 2, 8)
    assert m.encoder.embeddings.shape[0] == 6328848  # bound 2 table (cfg3/5)
    groups = m.get_params(1e-3)
    assert [g["lr"] for g in groups] == [1e-3, 1e-3, 1e-3, 2e-3] and groups[3]["params"] is m.color_palette
    assert m.get_params_but_dont_learn_palette(1e-3)[3]["lr"] == 0
    assert m.color_palette.shape == (8, 3) and m.color_palette.requires_grad


def test_reference_callers_construct_on_the_dropin_packages():
    """The reference's own nerf/network_ff.py + nerf/renderer.py (staged untouched by oracle/build_ref.py) import and construct with
    dropin/ in front of the path: `import raymarching`, `from gridencoder import GridEncoder`, `from ffmlp import FFMLP`, `from
    shencoder import SHEncoder` resolve to laenerf_b200 (the GPU tier then RUNS them: tests/test_gpu_refstack.py)."""
    import sys
    import pytest
    import ref_stack
    if not ref_stack.available("dropin"):
        pytest.skip("oracle/_ref/py not staged (python oracle/build_ref.py py)")
    before = {k: sys.modules.get(k) for k in ("raymarching", "gridencoder", "ffmlp", "shencoder", "nerf", "encoding")}
    m = ref_stack.make_model("dropin", "ff", bound=2, density_scale=1, min_near=0.2, density_thresh=10)
    assert type(m).__module__ == "nerf.network_ff" and type(m).__mro__[1].__module__ == "nerf.renderer"
    for sub in (m.encoder, m.sigma_net, m.color_net, m.encoder_dir):
        assert type(sub).__module__.startswith("laenerf_b200."), type(sub)
    assert m.cascade == 2 and m.density_bitfield.numel() == 2 * 128 ** 3 // 8
    assert [tuple(p.shape) for p in m.parameters()] == [(6328848, 2), (7168,), (11264,)]
    # the flavour's modules do not leak into (or replace anything in) the process-wide module table
    assert before == {k: sys.modules.get(k) for k in before}


def test_reference_round_sequence_from_death_histogram():
    """NeRFNetwork._reference_sequence reconstructs the n_step of every round of the reference's inference loop
    (renderer.py:353-379: n_step = clamp(N // n_alive, 1, 8), a ray dies in the round that offers it more samples than it has left)
    from the histogram of the rays' death sample indices -- checked against a literal emulation of that loop."""
    import numpy as np
    from laenerf_b200.nerf import NeRFNetwork
    rng = np.random.default_rng(0)
    for trial in range(40):
        N = int(rng.integers(50, 5000))
        max_steps = int(rng.choice([64, 256, 1024]))
        k = rng.integers(0, 300, size=N)
        k[rng.random(N) < 0.6] = 0
        hist = np.bincount(np.minimum(k, max_steps + 72), minlength=max_steps + 73)
        seq = NeRFNetwork._reference_sequence(hist, N, max_steps)
        alive, done, step, want = np.ones(N, bool), np.zeros(N, int), 0, []
        while step < max_steps:
            na = int(alive.sum())
            if na <= 0:
                break
            n = max(min(N // na, 8), 1)
            want.append(n)
            take = np.where(alive, np.minimum(n, k - done), 0)
            alive &= ~(alive & (take < n))
            done += take
            step += n
        assert seq == want, (trial, seq[:12], want[:12])
