"""GPU tier: the reference's OWN callers on this library, next to the same callers on the reference's extensions.

tests/ref_stack.py loads nerf/renderer.py + nerf/network_ff.py (+ encoding.py, activation.py) of the reference, staged untouched
under oracle/_ref/py by oracle/build_ref.py, twice: with the reference's wrapper packages on the reference's compiled extensions
("reference"), and with dropin/ in front of the path so that `import raymarching`, `from gridencoder import GridEncoder`, `from
ffmlp import FFMLP`, `from shencoder import SHEncoder` resolve to laenerf_b200 ("dropin").  Same weights, same rays:

  * run_cuda training branch: identical per-step sample counts, image / loss / gradients within fp16 tolerances;
  * run_cuda inference loop on a full 800 x 800 view: image-level parity -- PSNR of both renders against the training target
    within 0.05 dB of each other (BASELINE.json north_star), for the dropin stack AND for the product path
    (laenerf_b200.nerf.NeRFNetwork, fused kernels, device-driven rounds on the reference schedule);
  * run_cuda_distill against an edit grid;
  * update_extra_state (renderer.py:556-649) of the reference on its own extensions vs NeRFNetwork.update_extra_state (row f-2).
"""
import numpy as np
import pytest
import torch

from cases import scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def stacks():
    import ref_stack
    if not (ref_stack.available("reference") and ref_stack.available("dropin")):
        pytest.skip("oracle/_ref (extensions + staged python) not built")
    return ref_stack


def _target(ro, rd):
    """A smooth procedural 'photograph': colour as a function of where the ray enters the unit box (a stand-in for a dataset)."""
    p = ro + rd * ((ro.norm(dim=-1, keepdim=True) - 1.0).clamp(min=0.0))
    return (0.5 + 0.5 * torch.sin(3.0 * p + torch.tensor([0.0, 2.0, 4.0], device=p.device))).contiguous()


@pytest.fixture(scope="module")
def trained(dev):
    """laenerf_b200 model trained for a few hundred steps on the synthetic lego-shape scene so that images have structure."""
    from laenerf_b200.nerf import NeRFNetwork, TrainStep
    from laenerf_b200.scene import get_rays_np
    sc = scene("lego")
    torch.manual_seed(0)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_thresh=10.0).to(dev)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    step = TrainStep(m)
    rng = np.random.default_rng(0)
    for it in range(300):
        ro, rd, _ = get_rays_np(sc.poses[it % len(sc.poses)], sc.intrinsics, sc.H, sc.W, N=4096, rng=rng)
        ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
        step(ro, rd, _target(ro, rd))
        if it in (0, 16, 64):
            m.update_mean_count()
    step.optimizer.detach()  # hand the modules back to the plain path: state_dict() is the trained fp32 state
    m.mean_count = 0
    m.local_step = 0
    return sc, m


def _make(stacks, kind, sc, src, dev, variant="ff"):
    m = stacks.make_model(kind, variant, device=dev, bound=sc.bound, density_scale=1, min_near=sc.min_near, density_thresh=10.0)
    stacks.copy_state(m, src)
    m.mean_count, m.local_step = 0, 0
    return m


def _psnr(a, b):
    return float(-10.0 * torch.log10((a - b).square().mean()))


def test_reference_callers_train_branch_on_dropin(dev, stacks, trained):
    sc, ours = trained
    from laenerf_b200.scene import get_rays_np
    R, D = _make(stacks, "reference", sc, ours, dev), _make(stacks, "dropin", sc, ours, dev)
    assert type(D.encoder).__module__.startswith("laenerf_b200") and type(D.sigma_net).__module__.startswith("laenerf_b200")
    assert not type(R.encoder).__module__.startswith("laenerf_b200")
    ro, rd, _ = get_rays_np(sc.poses[1], sc.intrinsics, sc.H, sc.W, N=4096, rng=np.random.default_rng(4))
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = _target(ro, rd)
    outs = {}
    for name, m in (("R", R), ("D", D)):
        m.train()
        for it in range(2):  # step 0 sizes the sample buffer from the counter, step 1 runs with mean_count > 0 (renderer.py:643-647)
            m.zero_grad()
            torch.manual_seed(77 + it)  # march noise
            with torch.autocast("cuda", dtype=torch.float16):
                out = m.run_cuda(ro, rd, dt_gamma=0, bg_color=1, perturb=True, force_all_rays=False, max_steps=1024, T_thresh=1e-4)
                loss = torch.nn.functional.mse_loss(out["image"], gt, reduction="none").mean(-1).mean()
            (loss * 1024.0).backward()
            if it == 0:
                m.mean_count = int(m.step_counter[0, 0].item())
        outs[name] = dict(image=out["image"].detach(), depth=out["depth"].detach(), loss=float(loss), counter=m.step_counter[:2].clone(),
                          ge=m.encoder.embeddings.grad.detach().float() / 1024.0, gs=m.sigma_net.weights.grad.detach().float() / 1024.0,
                          gc=m.color_net.weights.grad.detach().float() / 1024.0)
    a, b = outs["D"], outs["R"]
    assert torch.equal(a["counter"], b["counter"])                     # sample counts of both steps: bit-exact
    assert int(a["counter"][1, 0]) > 100_000
    assert torch.allclose(a["image"], b["image"], rtol=0, atol=4e-3), float((a["image"] - b["image"]).abs().max())
    ok = torch.isfinite(b["depth"])
    assert torch.allclose(a["depth"][ok], b["depth"][ok], rtol=1e-3, atol=2e-3)
    assert abs(a["loss"] - b["loss"]) <= 2e-3 * abs(b["loss"]) + 1e-6
    for k, tol in (("ge", 0.15), ("gs", 0.08), ("gc", 0.08)):  # the reference accumulates MLP gradients in fp16 (2.5-12 % of max, cases.TOL)
        err = float((a[k] - b[k]).abs().max()) / float(b[k].abs().max())
        assert err <= tol, (k, err)


def test_full_view_image_parity_reference_vs_dropin_vs_product(dev, stacks, trained):
    """per-image PSNR within 0.05 dB (north_star): the 800 x 800 view through (1) the reference callers on the reference
    extensions, (2) the same callers on dropin/, (3) the product path on the reference round schedule."""
    sc, ours = trained
    from laenerf_b200.scene import get_rays_np
    R, D = _make(stacks, "reference", sc, ours, dev), _make(stacks, "dropin", sc, ours, dev)
    ro, rd, _ = get_rays_np(sc.poses[2], sc.intrinsics, sc.H, sc.W)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = _target(ro, rd)
    imgs = {}
    for name, m in (("R", R), ("D", D)):
        m.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            imgs[name] = m.run_cuda(ro, rd, dt_gamma=0, bg_color=1, perturb=False, max_steps=1024, T_thresh=1e-4, scale_depth=True)
    ours.eval()
    ours.fused, ours.device_loop, ours.render_schedule = True, True, "auto"
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        imgs["P"] = ours.render(ro, rd, perturb=False, bg_color=1, T_thresh=1e-4)
    psnr = {k: _psnr(v["image"], gt) for k, v in imgs.items()}
    assert 5.0 < psnr["R"] < 60.0, psnr  # a trained, non-trivial image
    assert abs(psnr["D"] - psnr["R"]) <= 0.05, psnr
    assert abs(psnr["P"] - psnr["R"]) <= 0.05, psnr
    # and directly against each other
    assert _psnr(imgs["D"]["image"], imgs["R"]["image"]) > 45.0
    assert _psnr(imgs["P"]["image"], imgs["R"]["image"]) > 45.0
    hit = imgs["R"]["image"].ne(1.0).any(-1)
    assert float(hit.float().mean()) > 0.1
    dR, dD, dP = imgs["R"]["depth"], imgs["D"]["depth"], imgs["P"]["depth"]
    ok = torch.isfinite(dR) & torch.isfinite(dD) & torch.isfinite(dP)
    assert float((dR[ok] - dD[ok]).abs().mean()) < 2e-3 and float((dR[ok] - dP[ok]).abs().mean()) < 2e-3


def test_run_cuda_distill_reference_vs_dropin_vs_product(dev, stacks, trained):
    sc, ours = trained
    from laenerf_b200.scene import get_rays_np
    R, D = _make(stacks, "reference", sc, ours, dev), _make(stacks, "dropin", sc, ours, dev)
    ro, rd, _ = get_rays_np(sc.poses[3], sc.intrinsics, sc.H, sc.W, inds=np.arange(200 * 800, 360 * 800))
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    edit = ours.density_bitfield.clone()
    edit[: edit.numel() // 2] = 0  # an "edit grid": the upper half (Morton order) of the occupied cells
    outs = {}
    for name, m in (("R", R), ("D", D), ("P", ours)):
        m.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs[name] = m.run_cuda_distill(ro, rd, edit, dt_gamma=0, bg_color=1, perturb=False, max_steps=1024, T_thresh=1e-4)
    r = outs["R"]
    assert float(r["weights_edit"].sum()) > 0
    for name in ("D", "P"):
        o = outs[name]
        assert set(r.keys()) <= set(o.keys())
        assert _psnr(o["image"], r["image"]) > 45.0
        for k in ("weights", "weights_edit"):
            assert float((o[k] - r[k]).abs().mean()) < 2e-3, (name, k)
        for k in ("depth", "depth_edit"):
            assert float((o[k] - r[k]).abs().mean()) < 5e-3, (name, k)
        assert float((o["x_term"] - r["x_term"]).abs().mean()) < 5e-3
        assert torch.equal(o["min_near"], r["min_near"])


@pytest.mark.parametrize("partial", [False, True])
def test_update_extra_state_against_the_reference_renderer(dev, stacks, trained, partial):
    """Row f-2 against the reference itself: NeRFRenderer.update_extra_state of the staged reference renderer on the reference's
    extensions vs NeRFNetwork.update_extra_state (csrc/occupancy.cu), same torch seed, same starting grid."""
    sc, ours = trained
    R = _make(stacks, "reference", sc, ours, dev)
    from laenerf_b200.nerf import NeRFNetwork
    P = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_thresh=10.0).to(dev)
    P.load_state_dict(ours.state_dict())
    it0 = 16 if partial else 0
    R.iter_density, P.iter_density = it0, it0
    R.local_step = P.local_step = 0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        torch.manual_seed(123)
        R.update_extra_state()
        torch.manual_seed(123)
        P.update_extra_state()
    gr, gp = R.density_grid, P.density_grid
    assert torch.equal(gr < 0, gp < 0)
    # same query points (same RNG stream); the densities differ by the fp16-accumulate (reference) vs fp32-accumulate MLP arithmetic
    rel = (gr - gp).abs() / gr.abs().clamp(min=1e-3)
    assert float(rel.mean()) < (2e-2 if partial else 5e-3), float(rel.mean())
    frac_off = float((rel > 0.1).float().mean())
    assert frac_off < (0.12 if partial else 1e-3), frac_off  # partial updates: duplicate draws make torch's own scatter order-dependent
    assert abs(R.mean_density - P.mean_density) <= 2e-2 * abs(R.mean_density) + 1e-6
    x = (R.density_bitfield ^ P.density_bitfield).int()
    mismatch = sum(int(((x >> k) & 1).sum()) for k in range(8)) / (x.numel() * 8)
    assert mismatch < (0.02 if partial else 2e-3), mismatch
    assert int(P.density_bitfield.count_nonzero()) > 0
