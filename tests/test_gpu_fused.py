"""GPU tier: the fused rows f-1 (NeRFNetwork.forward/backward as fused kernels) and f-4 (Adam + AMP glue) of SURVEY.md
section 8, checked against (a) the CPU oracle composed the way network_ff.py:51-79 composes the ops, (b) the
module-by-module CUDA path of the same package, (c) torch.optim.Adam + torch.amp.GradScaler."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from cases import scene, scene_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _h(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def _weights(rng, in_dim, nl, amp=0.25):
    return _h(rng.uniform(-amp, amp, 64 * (in_dim + 64 * (nl - 1) + 16)))


def _fused_fwd(dev, enc, dirs, ws, wc, ns, nc, ds, train):
    from laenerf_b200 import _native as N
    M = enc.shape[0]
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    enc_d, dirs_d, ws_d, wc_d = t(enc, torch.half), t(dirs, torch.float32), t(ws, torch.half), t(wc, torch.half)
    sig = torch.empty(M, device=dev)
    rgb = torch.empty(M, 3, device=dev)
    fb = torch.empty(ns + nc, M, 64, dtype=torch.half, device=dev) if train else None
    cin = torch.empty(M, 32, dtype=torch.half, device=dev) if train else None
    h0 = torch.empty(M, dtype=torch.half, device=dev) if train else None
    N.check(N.lib().lnrf_nerf_forward(N.ptr(enc_d), N.ptr(dirs_d), N.ptr(ws_d), N.ptr(wc_d), M, ns, nc, ds, int(train), N.ptr(fb),
                                      N.ptr(cin), N.ptr(h0), N.ptr(sig), N.ptr(rgb), N.stream()))
    torch.cuda.synchronize()
    return dict(sig=sig, rgb=rgb, fb=fb, cin=cin, h0=h0, enc=enc_d, ws=ws_d, wc=wc_d)


@pytest.mark.parametrize("M,ns,nc", [(128, 2, 3), (1280, 2, 3), (40064, 2, 3), (640, 3, 2)])
def test_fused_forward_matches_oracle_composition(dev, oracle_backend, M, ns, nc):
    """sigma_net -> trunc_exp -> SH -> cat -> color_net -> sigmoid restated with the CPU oracle's FFMLP / SH pieces."""
    rng = np.random.default_rng(M + ns)
    enc = _h(rng.standard_normal((M, 32)) * 0.5)
    dirs = rng.standard_normal((M, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    ws, wc = _weights(rng, 32, ns), _weights(rng, 32, nc)
    ds = 1.5
    got = _fused_fwd(dev, enc, dirs, ws, wc, ns, nc, ds, True)
    # oracle composition (fp16 storage points as in the reference: h, SH cast by custom_fwd(cast_inputs=half), rgb)
    h, fb_s = oracle_backend.ffmlp_fwd(enc, ws, 32, 16, 64, ns)
    h = _h(h)
    sig = ds * np.exp(h[:, 0].astype(np.float32))
    sh = _h(oracle_backend.sh(dirs, 4))
    cin = np.concatenate([sh, h[:, 1:], np.zeros((M, 1), np.float32)], axis=1)
    hc, fb_c = oracle_backend.ffmlp_fwd(cin, wc, 32, 16, 64, nc)
    rgb = _h(1.0 / (1.0 + np.exp(-_h(hc[:, :3]))))
    assert np.allclose(got["sig"].cpu().numpy(), sig, rtol=2e-2, atol=1e-3)      # exp amplifies the fp16 ulp of h0 (2^-11 rel. at |h0|~4 => ~1e-2)
    assert np.allclose(got["rgb"].cpu().numpy(), rgb, rtol=0, atol=2e-3)
    assert np.allclose(got["cin"].float().cpu().numpy()[:, :16], sh, rtol=0, atol=1e-3)
    assert np.allclose(got["cin"].float().cpu().numpy()[:, 16:], cin[:, 16:], rtol=2e-3, atol=2e-3)
    assert (got["cin"][:, 31] == 0).all()
    fb = got["fb"].float().cpu().numpy()
    assert np.allclose(fb[:ns], np.asarray(fb_s).reshape(ns, M, 64), rtol=2e-3, atol=2e-3)
    assert np.allclose(fb[ns:], np.asarray(fb_c).reshape(nc, M, 64), rtol=4e-3, atol=4e-3)
    # inference variant: same outputs, nothing saved
    inf = _fused_fwd(dev, enc, dirs, ws, wc, ns, nc, ds, False)
    assert torch.equal(inf["sig"], got["sig"]) and torch.equal(inf["rgb"], got["rgb"])


def _model(dev, fused, seed=0):
    from laenerf_b200.nerf import NeRFNetwork
    sc = scene("lego")
    torch.manual_seed(seed)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
    with torch.no_grad():  # make the table matter: the U(+-1e-4) init gives ~zero features
        m.encoder.embeddings.uniform_(-0.5, 0.5, generator=torch.Generator(device=dev).manual_seed(seed + 1))
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    m.fused = fused
    return m


def test_fused_network_equals_module_path_forward_and_backward(dev):
    """Same weights, same samples: the fused kernels against encoder -> FFMLP -> trunc_exp -> SH -> cat -> FFMLP -> sigmoid."""
    a, b = _model(dev, True), _model(dev, False)
    b.load_state_dict(a.state_dict())
    a.density_scale = b.density_scale = 1.0
    M = 128 * 37
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.rand(M, 3, device=dev, generator=g) * 1.6 - 0.8
    d = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1)
    gs = torch.randn(M, device=dev, generator=g) * 1e-2
    gr = torch.randn(M, 3, device=dev, generator=g) * 1e-2
    outs = []
    for m in (a, b):
        m.train()
        with torch.autocast("cuda", dtype=torch.float16):
            s, rgb = m.forward_scaled(x, d)
        assert s.shape == (M,) and rgb.shape == (M, 3)
        ((s.float() * gs).sum() * 64 + (rgb.float() * gr).sum() * 64).backward()
        outs.append((s.float().detach(), rgb.float().detach(), m.encoder.embeddings.grad, m.sigma_net.weights.grad, m.color_net.weights.grad))
    (s1, c1, ge1, gs1, gc1), (s2, c2, ge2, gs2, gc2) = outs
    assert torch.allclose(s1, s2, rtol=1e-5, atol=1e-6)            # same MMAs; expf vs torch.exp
    assert torch.allclose(c1, c2, rtol=0, atol=5e-4)               # one fp16 ulp of a sigmoid output
    for u, v, tol in ((ge1, ge2, 2e-2), (gs1, gs2, 1e-2), (gc1, gc2, 1e-2)):
        assert u is not None and v is not None
        assert float((u - v).abs().max()) <= tol * float(v.abs().max()) + 1e-7, float((u - v).abs().max() / v.abs().max())


def test_fused_network_renders_the_same_image(dev):
    a, b = _model(dev, True, 3), _model(dev, False, 3)
    b.load_state_dict(a.state_dict())
    _, ro, rd, _ = scene_rays("lego", 2048, 7)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    imgs = []
    for m in (a, b):
        m.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            imgs.append(m.render(ro, rd, perturb=False, bg_color=1)["image"])
    mse = float((imgs[0] - imgs[1]).square().mean())
    assert mse < 1e-7, mse  # PSNR between the two paths > 70 dB (the 0.05 dB bar is against ground truth)


def test_adam_step_matches_torch_adam(dev):
    """lnrf_adam_step on fp16 loss-scaled gradients against GradScaler.unscale_ + torch.optim.Adam (fused and foreach)."""
    from laenerf_b200 import _native as N
    n = 300_007  # ragged tail on purpose
    g = torch.Generator(device=dev).manual_seed(11)
    p0 = torch.randn(n, device=dev, generator=g) * 1e-2
    scale = 65536.0
    for impl in ("fused", "foreach"):
        p_ref = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, **({"fused": True} if impl == "fused" else {"foreach": True}))
        p, m, v = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        p16 = torch.empty(n, dtype=torch.half, device=dev)
        step = torch.ones(1, device=dev)
        found = torch.zeros(1, device=dev)
        sc = torch.full((1,), scale, device=dev)
        for it in range(4):
            g16 = (torch.randn(n, device=dev, generator=g) * 1e-3 * scale).half()
            g16[::7] = 0
            p_ref.grad = g16.float() / scale
            opt.step()
            gbuf = g16.clone()
            arr = (N.OptTensor * 1)()
            arr[0].params, arr[0].exp_avg, arr[0].exp_avg_sq, arr[0].grad = p.data_ptr(), m.data_ptr(), v.data_ptr(), gbuf.data_ptr()
            arr[0].params_f16, arr[0].n, arr[0].grad_dtype = p16.data_ptr(), n, N.F16
            N.check(N.lib().lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), 1, N.ptr(found), N.stream()))
            N.check(N.lib().lnrf_adam_step(C.cast(arr, C.c_void_p), 1, 1e-2, 0.9, 0.99, 1e-15, 0.0, N.ptr(sc), N.ptr(found), N.ptr(step),
                                           None, N.stream()))
            N.check(N.lib().lnrf_amp_update(N.ptr(sc), None, N.ptr(found), N.ptr(step), 2.0, 0.5, 2000, N.stream()))
            assert not gbuf.any(), "the gradient buffer must be cleared"
            assert float(step.item()) == it + 2
        st = opt.state[p_ref]
        tol = dict(rtol=2e-6, atol=1e-9)
        assert torch.allclose(m, st["exp_avg"], **tol)
        assert torch.allclose(v, st["exp_avg_sq"], rtol=2e-6, atol=1e-15)
        assert torch.allclose(p, p_ref.detach(), rtol=2e-6, atol=2e-8), float((p - p_ref.detach()).abs().max())
        assert torch.equal(p16, p.half())


def test_adam_step_skips_on_inf_like_gradscaler(dev):
    from laenerf_b200 import _native as N
    n = 4096 + 3
    p = torch.randn(n, device=dev)
    p_before = p.clone()
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    gbuf = torch.ones(n, dtype=torch.half, device=dev)
    gbuf[n - 2] = float("inf")
    step, found, sc, tracker = torch.ones(1, device=dev), torch.zeros(1, device=dev), torch.full((1,), 1024.0, device=dev), torch.full((1,), 5, dtype=torch.int32, device=dev)
    arr = (N.OptTensor * 1)()
    arr[0].params, arr[0].exp_avg, arr[0].exp_avg_sq, arr[0].grad = p.data_ptr(), m.data_ptr(), v.data_ptr(), gbuf.data_ptr()
    arr[0].params_f16, arr[0].n, arr[0].grad_dtype = None, n, N.F16
    N.check(N.lib().lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), 1, N.ptr(found), N.stream()))
    assert float(found.item()) == 1.0
    N.check(N.lib().lnrf_adam_step(C.cast(arr, C.c_void_p), 1, 1e-2, 0.9, 0.99, 1e-15, 0.0, N.ptr(sc), N.ptr(found), N.ptr(step), None, N.stream()))
    N.check(N.lib().lnrf_amp_update(N.ptr(sc), N.ptr(tracker), N.ptr(found), N.ptr(step), 2.0, 0.5, 2000, N.stream()))
    assert torch.equal(p, p_before) and not m.any() and not v.any() and not gbuf.any()
    assert float(sc.item()) == 512.0 and int(tracker.item()) == 0 and float(step.item()) == 1.0 and float(found.item()) == 0.0
    # growth after `growth_interval` clean steps
    tracker.fill_(1999)
    gbuf.fill_(1.0)
    N.check(N.lib().lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), 1, N.ptr(found), N.stream()))
    N.check(N.lib().lnrf_adam_step(C.cast(arr, C.c_void_p), 1, 1e-2, 0.9, 0.99, 1e-15, 0.0, N.ptr(sc), N.ptr(found), N.ptr(step), None, N.stream()))
    N.check(N.lib().lnrf_amp_update(N.ptr(sc), N.ptr(tracker), N.ptr(found), N.ptr(step), 2.0, 0.5, 2000, N.stream()))
    assert float(sc.item()) == 1024.0 and int(tracker.item()) == 0 and float(step.item()) == 2.0 and not torch.equal(p, p_before)


def test_one_launch_optimizer_step_equals_the_three_launch_sequence(dev):
    """lnrf_adam_amp_step (non-finite check + Adam + GradScaler.update in one launch, grid barrier inside) against
    lnrf_grad_nonfinite_check + lnrf_adam_step + lnrf_amp_update on the same data, including a skipped (inf) step and a scale growth:
    bit-identical parameters, moments, shadows, scale, tracker, step count."""
    from laenerf_b200 import _native as N
    lib = N.lib()
    sizes = [6119864 * 2, 7168, 11264]   # the lego-shape table and the two MLPs
    g = torch.Generator(device=dev).manual_seed(21)

    def fresh():
        st = []
        g2 = torch.Generator(device=dev).manual_seed(22)
        for n in sizes:
            p = torch.randn(n, device=dev, generator=g2) * 1e-2
            st.append(dict(p=p, m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev), p16=p.half(), g=torch.zeros(n, dtype=torch.half, device=dev)))
        return dict(t=st, step=torch.ones(1, device=dev), found=torch.zeros(1, device=dev), scale=torch.full((1,), 8192.0, device=dev),
                    tracker=torch.full((1,), 1, dtype=torch.int32, device=dev), sync=torch.zeros(4, dtype=torch.int32, device=dev))

    def desc(S):
        arr = (N.OptTensor * len(sizes))()
        for i, t in enumerate(S["t"]):
            arr[i].params, arr[i].exp_avg, arr[i].exp_avg_sq, arr[i].grad = t["p"].data_ptr(), t["m"].data_ptr(), t["v"].data_ptr(), t["g"].data_ptr()
            arr[i].params_f16, arr[i].n, arr[i].grad_dtype = t["p16"].data_ptr(), sizes[i], N.F16
        return arr

    A, B = fresh(), fresh()
    hyper = (1e-2, 0.9, 0.99, 1e-15, 0.0)
    for it in range(4):
        grads = [(torch.randn(n, device=dev, generator=g) * 4.0).half() for n in sizes]
        if it == 1:
            grads[0][123457] = float("nan")   # this step must be skipped
        for S in (A, B):
            for t, gr in zip(S["t"], grads):
                t["g"].copy_(gr)
        arr = desc(A)
        N.check(lib.lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), 3, N.ptr(A["found"]), None))
        N.check(lib.lnrf_adam_step(C.cast(arr, C.c_void_p), 3, *hyper, N.ptr(A["scale"]), N.ptr(A["found"]), N.ptr(A["step"]), None, None))
        N.check(lib.lnrf_amp_update(N.ptr(A["scale"]), N.ptr(A["tracker"]), N.ptr(A["found"]), N.ptr(A["step"]), 2.0, 0.5, 3, None))
        arr = desc(B)
        N.check(lib.lnrf_adam_amp_step(C.cast(arr, C.c_void_p), 3, *hyper, N.ptr(B["scale"]), N.ptr(B["tracker"]), N.ptr(B["found"]), N.ptr(B["step"]),
                                       None, 2.0, 0.5, 3, N.ptr(B["sync"]), None))
        torch.cuda.synchronize()
        for ta, tb in zip(A["t"], B["t"]):
            for k in ("p", "m", "v", "p16", "g"):
                assert torch.equal(ta[k], tb[k]), (it, k)
            assert not tb["g"].any()
        for k in ("step", "found", "scale", "tracker"):
            assert torch.equal(A[k], B[k]), (it, k, A[k], B[k])
        assert B["sync"][0].item() == 0 and B["sync"][2].item() == 0 and B["sync"][1].item() == it + 1
    # steps 0, 2, 3 ran, step 1 was skipped: scale backed off once (8192 -> 4096), two clean steps since (growth needs three)
    assert float(B["step"]) == 4.0 and float(B["scale"]) == 4096.0 and int(B["tracker"]) == 2


def test_train_step_fused_optimizer_tracks_torch_path(dev):
    """Whole training steps: AmpAdam + fused network against torch Adam/GradScaler + module path from the same init.
    Hash-grid atomics make both paths run-to-run nondeterministic in the last fp16 bits, so the check is statistical:
    the losses agree and the parameters stay close after several steps."""
    from laenerf_b200.nerf import TrainStep
    a, b = _model(dev, True, 9), _model(dev, False, 9)
    b.load_state_dict(a.state_dict())
    sa, sb = TrainStep(a, fused_optimizer=True), TrainStep(b, fused_optimizer=False)
    assert sa.fused_optimizer and not sb.fused_optimizer
    _, ro, rd, _ = scene_rays("lego", 2048, 13)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = torch.rand(2048, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    la, lb = [], []
    for it in range(6):
        torch.manual_seed(100 + it)  # same march noise on both paths
        la.append(float(sa(ro, rd, gt)[0]))
        torch.manual_seed(100 + it)
        lb.append(float(sb(ro, rd, gt)[0]))
    assert all(np.isfinite(la)) and la[-1] < la[0]
    assert np.allclose(la, lb, rtol=2e-2), (la, lb)
    assert a.encoder.embeddings.grad is None and a.sigma_net.weights.grad is None  # gradients never materialise in fp32
    for pa, pb in ((a.sigma_net.weights, b.sigma_net.weights), (a.color_net.weights, b.color_net.weights)):
        assert float((pa - pb).abs().max()) < 0.08  # 6 Adam steps of lr 1e-2 move a weight by <= 0.06
    assert torch.equal(a.encoder._shadow_f16, a.encoder.embeddings.detach().half())
    sd = sa.optimizer.state_dict()
    assert float(sd["state"][0]["step"]) == 6 and sd["state"][0]["exp_avg"].shape == a.encoder.embeddings.shape


def test_amp_adam_checkpoints_interchange_with_torch_adam(dev):
    """state_dict() has torch.optim.Adam's layout with the reference's four parameter groups (network_ff.py:139-153: encoder,
    sigma_net, an EMPTY encoder_dir group, color_net): torch's optimizer built from model.get_params() loads it, and AmpAdam
    loads torch's; the GradScaler state is separate, in torch.amp.GradScaler's layout."""
    from laenerf_b200.nerf import TrainStep
    a = _model(dev, True, 51)
    sa = TrainStep(a)
    _, ro, rd, _ = scene_rays("lego", 1024, 3)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = torch.rand(1024, 3, device=dev)
    for _ in range(3):
        sa(ro, rd, gt)
    sd = sa.optimizer.state_dict()
    assert [g["params"] for g in sd["param_groups"]] == [[0], [1], [], [2]] and "scaler" not in sd
    topt = torch.optim.Adam(a.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    topt.load_state_dict(sd)                                   # AmpAdam -> torch
    st = topt.state[a.encoder.embeddings]
    assert float(st["step"]) == 3 and torch.equal(st["exp_avg"], sa.optimizer.state[0]["exp_avg"])
    scaler = torch.amp.GradScaler("cuda")
    scaler.load_state_dict(sa.optimizer.scaler_state_dict())   # scaler state in GradScaler's own layout
    assert scaler.get_scale() == sa.optimizer.get_scale()
    b = _model(dev, True, 51)
    sb = TrainStep(b)
    sb.optimizer.load_state_dict(topt.state_dict())            # torch -> AmpAdam
    sb.optimizer.load_scaler_state_dict(scaler.state_dict())
    assert float(sb.optimizer.step_count) == 4.0
    for x, y in zip(sb.optimizer.state, sa.optimizer.state):
        assert torch.equal(x["exp_avg"], y["exp_avg"]) and torch.equal(x["exp_avg_sq"], y["exp_avg_sq"])


def test_fp16_shadows_follow_outside_writes_to_the_fp32_parameters(dev):
    """the kernels read AmpAdam's fp16 shadows; load_state_dict / reset_parameters / an EMA copy_to() write the fp32
    parameters behind its back.  torch bumps `_version` on those writes, the shadow is re-derived before its next use
    (laenerf_b200/_shadow.py) -- the reference re-casts on every forward (grid.py:43-44)."""
    from laenerf_b200.nerf import TrainStep
    a, donor = _model(dev, True, 61), _model(dev, True, 62)
    with torch.no_grad():
        donor.sigma_net.weights.mul_(0.5)
    sa = TrainStep(a)
    x = torch.rand(1280, 3, device=dev) * 1.6 - 0.8
    d = torch.nn.functional.normalize(torch.randn(1280, 3, device=dev), dim=-1)
    a.eval()

    def out():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return [t.clone() for t in a.forward_scaled(x, d)]

    before = out()
    donor.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        want = [t.clone() for t in donor.forward_scaled(x, d)]
    assert not torch.equal(before[0], want[0])
    a.load_state_dict(donor.state_dict())          # after AmpAdam construction
    after = out()
    assert torch.equal(after[0], want[0]) and torch.equal(after[1], want[1])
    assert torch.equal(a.encoder._shadow_f16, a.encoder.embeddings.detach().half())
    # and the optimizer keeps training from the loaded state (its own raw-pointer writes do not trigger a re-derivation)
    _, ro, rd, _ = scene_rays("lego", 1024, 5)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    sa(ro, rd, torch.rand(1024, 3, device=dev))
    assert torch.equal(a.encoder._shadow_f16, a.encoder.embeddings.detach().half())
    assert not torch.equal(a.sigma_net.weights.detach(), donor.sigma_net.weights.detach())


def test_graphed_train_step_with_fused_optimizer(dev):
    from laenerf_b200.nerf import GraphedTrainStep, TrainStep
    m = _model(dev, True, 21)
    step = TrainStep(m)
    _, ro, rd, _ = scene_rays("lego", 4096, 17)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = torch.rand(4096, 3, device=dev)
    gs = GraphedTrainStep(step, 4096)
    losses = [float(gs(ro, rd, gt)[0]) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert float(step.optimizer.step_count.item()) >= 9  # warm-up + capture + replays all advanced the device-side step number


def test_whole_training_steps_against_the_reference_extensions(dev):
    """Several full training steps (march -> encode -> MLPs -> composite -> loss -> backward -> Adam) through the
    reference's own extensions (tests/ref_step.py on oracle/_ref) and through laenerf_b200 (fused path) from the same
    initial state and the same march noise: sample counts are identical, losses agree to fp16-training accuracy."""
    import ref_step
    if not ref_step.ref_available():
        pytest.skip("oracle/_ref not built")
    from laenerf_b200.nerf import TrainStep
    ours = _model(dev, True, 31)
    ref = ref_step.RefNeRF(ours).to(dev)
    so, sr = TrainStep(ours), ref_step.RefTrainStep(ref)
    _, ro, rd, _ = scene_rays("lego", 2048, 19)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = torch.rand(2048, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    lo, lr = [], []
    for it in range(5):
        torch.manual_seed(200 + it)
        l, out = so(ro, rd, gt)
        lo.append(float(l))
        torch.manual_seed(200 + it)
        l2, m2 = sr(ro, rd, gt)
        lr.append(float(l2))
        assert int(ours.step_counter[(ours.local_step - 1) % 16, 0]) == int(ref.step_counter[(ref.local_step - 1) % 16, 0])
    assert np.allclose(lo, lr, rtol=2e-2), (lo, lr)
    assert lo[-1] < lo[0] and lr[-1] < lr[0]


def test_device_driven_render_equals_host_loop(dev):
    """Row f-3: the inference rounds driven from the device (csrc/render.cu) against the host loop of run_cuda -- same kernels,
    same per-round geometry, so image / depth / weights are bit-identical; also through the distillation variant."""
    m = _model(dev, True, 41)
    _, ro, rd, _ = scene_rays("lego", 6000, 23)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    m.eval()
    m.render_schedule = "reference"   # the host loop IS the reference's n_step rule: same rounds, same slot counts
    outs = []
    for loop in (True, False):
        m.device_loop = loop
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs.append(m.render(ro, rd, perturb=False, bg_color=1))
    a, b = outs
    assert a["rounds"] > 3 and a["num_points"] == b["num_points"]
    assert torch.equal(a["image"], b["image"])
    assert torch.allclose(a["depth"], b["depth"], rtol=0, atol=0, equal_nan=True)  # rays that miss the box: 0 / (far - near = 0)
    # distillation render with a synthetic edit grid (every other byte of the density bitfield)
    edit = m.density_bitfield.clone()
    edit[::2] = 0
    outs = []
    for loop in (True, False):
        m.device_loop = loop
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs.append(m.run_cuda_distill(ro, rd, edit, perturb=False))
    a, b = outs
    for k in ("image", "depth", "weights_sum", "weights_edit_sum", "depth_edit"):
        assert torch.allclose(a[k], b[k], rtol=0, atol=0, equal_nan=True), k
    assert float(a["weights_edit_sum"].sum()) > 0
    m.device_loop = True
    m.render_schedule = "auto"


def test_grid_encoder_world_coordinates_equal_prenormalised_inputs(dev):
    from laenerf_b200.gridencoder import GridEncoder, grid_encode
    enc = GridEncoder(desired_resolution=4096).to(dev)
    with torch.no_grad():
        enc.embeddings.uniform_(-1, 1)
    x = torch.rand(5000, 3, device=dev) * 4 - 2
    for bound in (1, 2, 16):
        with torch.autocast("cuda", dtype=torch.float16):
            y = enc(x.clamp(-bound, bound), bound=bound)
            ref = grid_encode((x.clamp(-bound, bound) + bound) / (2 * bound), enc.embeddings, enc.offsets, enc.per_level_scale,
                              enc.base_resolution, False, enc.gridtype_id, enc.align_corners, enc.interp_id)
        assert torch.equal(y, ref)


def _update_extra_state_torch(m, decay=0.95):
    """Plain-torch restatement of NeRFRenderer.update_extra_state (renderer.py:556-649) on the module path of the package:
    the comparison target for the fused row f-2 path (same RNG calls in the same order)."""
    from laenerf_b200 import raymarching
    H, dev = m.grid_size, m.density_bitfield.device
    tmp_grid = -torch.ones_like(m.density_grid)
    if m.iter_density < 16:
        ar = torch.arange(H, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
        indices = raymarching.morton3D(coords).long()
        xyzs = 2 * coords.float() / (H - 1) - 1
        for cas in range(m.cascade):
            bound = min(2 ** cas, m.bound)
            hgs = bound / H
            cas_xyzs = xyzs * (bound - hgs)
            cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * hgs
            sigmas = m.density(cas_xyzs)["sigma"].reshape(-1).detach().float()
            sigmas *= m.density_scale
            tmp_grid[cas, indices] = sigmas
    else:
        n4 = H ** 3 // 4
        for cas in range(m.cascade):
            coords = torch.randint(0, H, (n4, 3), device=dev)
            indices = raymarching.morton3D(coords).long()
            occ_indices = torch.nonzero(m.density_grid[cas] > 0).squeeze(-1)
            rand_mask = torch.randint(0, occ_indices.shape[0], [n4], dtype=torch.long, device=dev)
            occ_indices = occ_indices[rand_mask]
            occ_coords = raymarching.morton3D_invert(occ_indices)
            indices = torch.cat([indices, occ_indices], dim=0)
            coords = torch.cat([coords, occ_coords], dim=0)
            xyzs = 2 * coords.float() / (H - 1) - 1
            bound = min(2 ** cas, m.bound)
            hgs = bound / H
            cas_xyzs = xyzs * (bound - hgs)
            cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * hgs
            sigmas = m.density(cas_xyzs)["sigma"].reshape(-1).detach().float()
            sigmas *= m.density_scale
            tmp_grid[cas, indices] = sigmas
    valid = (m.density_grid >= 0) & (tmp_grid >= 0)
    m.density_grid[valid] = torch.maximum(m.density_grid[valid] * decay, tmp_grid[valid])
    m.mean_density = torch.mean(m.density_grid.clamp(min=0)).item()
    m.iter_density += 1
    raymarching.packbits(m.density_grid, min(m.mean_density, m.density_thresh), m.density_bitfield)


@pytest.mark.parametrize("bound", [1, 2])
def test_update_extra_state_matches_torch_restatement(dev, bound):
    """Row f-2: the fused occupancy update against the torch restatement of the reference's, same seed."""
    from laenerf_b200.nerf import NeRFNetwork
    torch.manual_seed(7)
    a = NeRFNetwork(bound=bound, density_thresh=0.01).to(dev)
    with torch.no_grad():
        a.encoder.embeddings.uniform_(-0.5, 0.5)
    b = NeRFNetwork(bound=bound, density_thresh=0.01).to(dev)
    b.load_state_dict(a.state_dict())
    b.fused = False
    nbits = a.density_bitfield.numel() * 8

    def bit_mismatch():
        x = (a.density_bitfield ^ b.density_bitfield).int()
        return sum(int(((x >> k) & 1).sum()) for k in range(8)) / nbits

    for it in range(2):  # two full updates (the second exercises the EMA-max)
        with torch.autocast("cuda", dtype=torch.float16):
            torch.manual_seed(100 + it)
            a.update_extra_state()
            torch.manual_seed(100 + it)
            _update_extra_state_torch(b)
        assert torch.equal(a.density_grid, b.density_grid), float((a.density_grid - b.density_grid).abs().max())  # same RNG, same roundings
        assert abs(a.mean_density - b.mean_density) <= 1e-3 * abs(b.mean_density) + 1e-6
        assert bit_mismatch() < 1e-3
    assert a.iter_density == 2 and int(a.density_bitfield.count_nonzero()) > 0
    # partial update: duplicates make the indexed assignment order-dependent in torch too, so the check is statistical
    a.iter_density = b.iter_density = 16
    with torch.autocast("cuda", dtype=torch.float16):
        torch.manual_seed(300)
        a.update_extra_state()
        torch.manual_seed(300)
        _update_extra_state_torch(b)
    changed_a, changed_b = (a._tmp_grid == -1).all(), True
    assert changed_a and changed_b  # the scratch grid is re-armed for the next update
    assert abs(a.mean_density - b.mean_density) <= 2e-2 * abs(b.mean_density) + 1e-6
    # ~9 % of the cells are drawn more than once (1 M draws over 2 M cells) and either draw may win the assignment
    assert (a.density_grid - b.density_grid).abs().gt(1e-3 + 5e-3 * b.density_grid.abs()).float().mean() < 0.10
    assert bit_mismatch() < 0.02


def test_lookahead_graph_trains_the_same_sequence(dev):
    """GraphedTrainStep(lookahead=True) marches batch k beside the backward of batch k-1 (two sample-buffer sets, two graphs): same
    batches, same updates, the loss simply arrives one call later.  Half-way the occupancy bitfield changes (what update_extra_state
    does every 16 steps): `remarch()` sends the batch in flight through the new bitfield, as the un-pipelined loop would have."""
    from laenerf_b200.nerf import GraphedTrainStep, TrainStep
    batches = []
    for k in range(7):
        _, ro, rd, _ = scene_rays("lego", 4096, 50 + k)
        batches.append((torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev),
                        torch.rand(4096, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(k))))
    runs, counts = [], []
    for look in (False, True):
        m = _model(dev, True, 61)
        g = GraphedTrainStep(TrainStep(m), 4096, perturb=False, lookahead=look)
        g.capture(*batches[0], warmup=1)
        losses, pts = [], []

        def change_grid():
            bf = m.density_bitfield
            bf[: bf.numel() // 3] = 0
            g.remarch()

        if look:
            # call k trains on batch k-1 (batch 0 was primed by capture); the loss is a view of graph-owned memory that the next
            # replay overwrites, so it is read before the next call
            for k, b in enumerate(batches[1:], 1):
                losses.append(float(g(*b)[0]))
                pts.append(int(g.sets[g.phase]["counter"][0]))   # samples of the batch just marched (batch k)
                if k == 3:
                    change_grid()   # batch 3 is in flight: marched through the old grid, must be re-marched
                    pts[-1] = int(g.sets[g.phase]["counter"][0])
            losses.append(float(g.flush()[0]))
        else:
            for k, b in enumerate(batches):
                if k == 3:
                    change_grid()
                losses.append(float(g(*b)[0]))
                if k:
                    pts.append(int(m.step_counter[0, 0]))
        runs.append(losses)
        counts.append(pts)
    seq, pipe = runs
    assert counts[0] == counts[1], counts   # identical sample counts per batch, also across the grid change
    assert counts[0][2] < counts[0][1] * 0.95   # ... which really removed samples
    assert len(seq) == len(pipe) == 7 and all(np.isfinite(seq)) and all(np.isfinite(pipe))
    assert np.allclose(seq, pipe, rtol=2e-2), (seq, pipe)


@pytest.mark.parametrize("bg_kind,scale_depth", [("scalar", True), ("per_ray", True), ("scalar", False)])
def test_composite_loss_tail_matches_module_tier_and_torch_autograd(dev, bg_kind, scale_depth):
    """Row f-5: composite + blend + depth normalisation + MSE in one launch each way against composite_rays_train followed by
    the torch expressions of renderer.py:324-329 / nerf/utils.py:592,633 and torch autograd (fp32; differences are summation
    order only)."""
    from laenerf_b200 import raymarching
    from laenerf_b200.nerf import NeRFNetwork
    sc = scene("lego")
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    m.train()
    n = 1500  # ragged: not a multiple of the 8 rays a block composites
    _, ro, rd, _ = scene_rays("lego", n, 23)
    torch.manual_seed(5)
    mr = m.march_train(torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev), perturb=True)
    M = mr["xyzs"].shape[0]
    g = torch.Generator(device=dev).manual_seed(3)
    sig0 = torch.rand(M, device=dev, generator=g) * 40.0  # large enough that some rays hit the T_thresh early-out
    rgb0 = torch.rand(M, 3, device=dev, generator=g)
    gt = torch.rand(n, 3, device=dev, generator=g)
    bg = torch.rand(n, 3, device=dev, generator=g) if bg_kind == "per_ray" else 1
    scale = torch.tensor(1024.0, device=dev)

    sa, ra = sig0.clone().requires_grad_(True), rgb0.clone().requires_grad_(True)
    ws, depth, image = raymarching.composite_rays_train(sa, ra, mr["deltas"], mr["rays"], 1e-4)
    image = image + (1 - ws).unsqueeze(-1) * bg
    if scale_depth:
        depth = torch.clamp(depth - mr["nears"], min=0) / (mr["fars"] - mr["nears"])
    loss_a = torch.nn.functional.mse_loss(image, gt, reduction="none").mean(-1).mean()
    (loss_a * scale).backward()

    sb, rb = sig0.clone().requires_grad_(True), rgb0.clone().requires_grad_(True)
    loss_b, ws_b, depth_b, image_b = raymarching.composite_loss_train(sb, rb, mr["deltas"], mr["rays"], gt, bg,
                                                                      mr["nears"] if scale_depth else None,
                                                                      mr["fars"] if scale_depth else None, 1e-4)
    torch.autograd.backward(loss_b, grad_tensors=scale)
    torch.cuda.synchronize()
    assert torch.equal(ws, ws_b)  # same kernel body
    assert float((image - image_b).abs().max()) <= 1e-6
    # rays that miss the box have far == near: 0 / 0 on both paths (renderer.py:328 does the same)
    assert torch.allclose(depth, depth_b, rtol=0, atol=1e-6, equal_nan=True) and torch.equal(depth.isnan(), depth_b.isnan())
    assert abs(float(loss_a) - float(loss_b)) <= 2e-6 * abs(float(loss_a))
    for x, y in ((sa.grad, sb.grad), (ra.grad, rb.grad)):
        tol = 1e-5 * float(x.abs().max())
        assert float((x - y).abs().max()) <= tol, (float((x - y).abs().max()), tol)
    assert float(sb.grad.abs().max()) > 0 and not ws_b.requires_grad and not image_b.requires_grad
    # deterministic: a second launch reproduces the loss bit for bit (fixed-order reduction, re-armed ticket)
    loss_c = raymarching.composite_loss_train(sb.detach(), rb.detach(), mr["deltas"], mr["rays"], gt, bg, None, None, 1e-4)[0]
    loss_d = raymarching.composite_loss_train(sb.detach(), rb.detach(), mr["deltas"], mr["rays"], gt, bg, None, None, 1e-4)[0]
    assert float(loss_c) == float(loss_d) == float(loss_b)


@pytest.mark.parametrize("bg_kind,scale_depth,short_buffer", [("scalar", True, False), ("per_ray", False, False), ("scalar", True, True)])
def test_composite_loss_forward_backward_in_one_launch_is_bit_identical(dev, bg_kind, scale_depth, short_buffer):
    """lnrf_composite_loss_train_forward_backward (the AMP scale handed to the forward) against the two launches: loss, images,
    depth and both sample gradients bit for bit -- also with a sample buffer too short for the last rays (dropped rays, zero-filled
    gradient rows) -- and the fall-back when autograd delivers a different gradient than the forward was promised."""
    from laenerf_b200 import _native as N
    from laenerf_b200 import raymarching
    from laenerf_b200.nerf import NeRFNetwork
    sc = scene("lego")
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    m.train()
    n = 1500
    _, ro, rd, _ = scene_rays("lego", n, 23)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    torch.manual_seed(5)
    mr = m.march_train(ro, rd, perturb=True)
    if short_buffer:
        m.mean_count = int(mr["counter"][0]) * 2 // 3
        torch.manual_seed(5)
        mr = m.march_train(ro, rd, perturb=True)
        assert int(mr["counter"][0]) > mr["xyzs"].shape[0]   # rays were dropped
    M = mr["xyzs"].shape[0]
    g = torch.Generator(device=dev).manual_seed(3)
    sig0 = torch.rand(M, device=dev, generator=g) * 40.0
    rgb0 = torch.rand(M, 3, device=dev, generator=g)
    gt = torch.rand(n, 3, device=dev, generator=g)
    bg = torch.rand(n, 3, device=dev, generator=g) if bg_kind == "per_ray" else 1
    scale = torch.tensor(1024.0, device=dev)
    nf = (mr["nears"], mr["fars"]) if scale_depth else (None, None)
    runs = []
    for promised, delivered in ((None, scale), (scale, scale.view(())), (scale, torch.tensor(512.0, device=dev))):
        s_, r_ = sig0.clone().requires_grad_(True), rgb0.clone().requires_grad_(True)
        l0 = N.launch_count()
        out = raymarching.composite_loss_train(s_, r_, mr["deltas"], mr["rays"], gt, bg, nf[0], nf[1], 1e-4, promised)
        torch.autograd.backward(out[0], grad_tensors=delivered)
        runs.append((out, s_.grad, r_.grad, N.launch_count() - l0))
    (oa, gsa, gra, la), (ob, gsb, grb, lb), (oc, gsc, grc, lc) = runs
    assert (la, lb, lc) == (2, 1, 2)   # two launches / one / one + the fall-back
    for x, y in zip(oa, ob):
        assert torch.equal(x, y) or (torch.equal(x.isnan(), y.isnan()) and torch.equal(x.nan_to_num(), y.nan_to_num()))
    assert torch.equal(gsa, gsb) and torch.equal(gra, grb) and float(gsa.abs().max()) > 0
    # the fall-back really used the delivered gradient: exactly half (a power of two) of the other runs'
    assert torch.equal(gsc * 2, gsa) and torch.equal(grc * 2, gra)


def test_train_step_fused_loss_equals_unfused_loss_path(dev):
    """TrainStep(fused_loss=True) and (fused_loss=False) from the same state: same sample counts, losses equal to fp32
    summation order on the first step and close afterwards (hash-grid atomics reorder fp16 additions)."""
    from laenerf_b200.nerf import TrainStep
    a, b = _model(dev, True, 41), _model(dev, True, 41)
    b.load_state_dict(a.state_dict())
    sa, sb = TrainStep(a, fused_loss=True), TrainStep(b, fused_loss=False)
    _, ro, rd, _ = scene_rays("lego", 2048, 29)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    gt = torch.rand(2048, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
    la, lb = [], []
    for it in range(5):
        torch.manual_seed(300 + it)
        l, oa = sa(ro, rd, gt)
        la.append(float(l))
        torch.manual_seed(300 + it)
        l, ob = sb(ro, rd, gt)
        lb.append(float(l))
        assert oa["num_points"] == ob["num_points"]
    assert abs(la[0] - lb[0]) <= 1e-5 * abs(lb[0]), (la, lb)
    assert np.allclose(la, lb, rtol=1e-2), (la, lb)
    assert la[-1] < la[0]


def test_fast_render_schedule_is_bit_identical_where_the_marcher_proves_it(dev):
    """The reference's n_step rule only decides WHERE a ray's t passes through rays_t; composite_rays rebuilds rays_t from the
    deltas (raymarching.cu:1006), and whenever every delta is exactly representable the rebuilt value is the marcher's own t, so
    the round boundaries leave no trace.  The marcher checks that per sample (ctl[12], csrc/raymarch.cu):
      * lego shape (cameras outside the box, t >= 2): no flag, "fast" (far fewer rounds) == "reference" bit for bit -- image,
        depth, weights, and the distillation outputs;
      * bonsai / flower shapes (cameras inside the volume): the flag is raised for ~1 % of the rays; "auto" re-renders those rays on
        the reference's own n_step sequence and is again the reference schedule bit for bit."""
    from laenerf_b200.nerf import NeRFNetwork
    m = _model(dev, True, 43)
    m.density_scale = 20.0  # dense enough for the T_thresh early-out to kill rays inside a round
    _, ro, rd, _ = scene_rays("lego", 8192, 27)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    m.eval()
    outs = {}
    for sched in ("reference", "fast", "auto"):
        m.render_schedule = sched
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs[sched] = m.render(ro, rd, perturb=False, bg_color=1, scale_depth=False)
    a, b, c = outs["reference"], outs["fast"], outs["auto"]
    assert b["rounds"] * 3 < a["rounds"] * 2, (a["rounds"], b["rounds"])
    assert c["schedule"].startswith("fast") and c["rounds"] == b["rounds"]
    for k in ("image", "depth", "t"):
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k
    edit = m.density_bitfield.clone()
    edit[::2] = 0
    d = {}
    for sched in ("reference", "fast"):
        m.render_schedule = sched
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            d[sched] = m.run_cuda_distill(ro, rd, edit, perturb=False)
    for k in ("image", "weights", "weights_edit", "depth", "depth_edit"):
        assert torch.equal(d["reference"][k], d["fast"][k]), k
    # scenes where deltas are NOT all exact (cameras inside the volume): the flag comes up, the flagged rays are re-rendered on the
    # reference's n_step sequence (reconstructed from the histogram of the rays' death samples) and "auto" is still the reference
    # schedule bit for bit -- plain render and distillation render
    for name, n in (("bonsai", 40000), ("flower", 30000)):
        sc = scene(name)
        torch.manual_seed(3)
        mb = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=5.0).to(dev)
        with torch.no_grad():
            mb.encoder.embeddings.uniform_(-0.5, 0.5)
        mb.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
        _, ro, rd, _ = scene_rays(name, n, 29)
        ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
        mb.eval()
        outs = {}
        for sched in ("reference", "fast", "auto"):
            mb.render_schedule = sched
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                outs[sched] = mb.render(ro, rd, perturb=False, bg_color=1, scale_depth=False)
        assert outs["auto"]["schedule"].startswith("fast + ") and "one pass" in outs["auto"]["schedule"], outs["auto"]["schedule"]
        if name == "bonsai":   # the same fix-up run round by round through the device loop (lnrf_render_rounds with nstep_seq)
            os.environ["LNRF_FIXUP_ROUNDS"] = "1"
            try:
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                    by_rounds = mb.render(ro, rd, perturb=False, bg_color=1, scale_depth=False)
            finally:
                del os.environ["LNRF_FIXUP_ROUNDS"]
            assert "one pass" not in by_rounds["schedule"] and by_rounds["schedule"].startswith("fast + ")
            for k in ("image", "depth", "t"):
                assert torch.equal(outs["reference"][k], by_rounds[k]), (name, "by rounds", k)
        assert not torch.equal(outs["reference"]["image"], outs["fast"]["image"])      # the fast schedule ALONE is not exact here
        for k in ("image", "depth", "t"):
            assert torch.equal(outs["reference"][k], outs["auto"][k]), (name, k)
        edit = mb.density_bitfield.clone()
        edit[::2] = 0
        d = {}
        for sched in ("reference", "auto"):
            mb.render_schedule = sched
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                d[sched] = mb.run_cuda_distill(ro, rd, edit, perturb=False)
        for k in ("image", "weights", "weights_edit", "depth", "depth_edit"):
            assert torch.equal(d["reference"][k], d["auto"][k]), (name, "distill", k)


@pytest.mark.parametrize("name,n", [("lego", 60000), ("bonsai", 40000)])
def test_render_economies_change_no_bit(dev, name, n):
    """The exact economies of the device-driven fast rounds, each switched off in turn, against all of them on:
    rays clipped to the box around the occupied cells (lnrf_render_desc.occupied_box), compact rounds (more than 8 samples per ray
    per round: samples back to back, 4-lane compositor) against the reference's slot layout (8 samples per round), and the box
    itself: it must contain every occupied cell of every cascade."""
    from laenerf_b200 import raymarching
    from laenerf_b200.nerf import NeRFNetwork
    sc = scene(name)
    torch.manual_seed(7)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=5.0).to(dev)
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.5, 0.5)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    _, ro, rd, _ = scene_rays(name, n, 31)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    m.eval()
    m.render_schedule = "auto"   # 8 and 64 samples per round are different schedules: only "auto" makes both the reference's bits
    edit = m.density_bitfield.clone()
    edit[1::2] = 0

    def run():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            a = m.render(ro, rd, perturb=False, bg_color=1, scale_depth=False)
            b = m.run_cuda_distill(ro, rd, edit, perturb=False)
        return a, b

    base = run()
    assert base[0]["rounds"] < 40
    variants = {}
    m.render_clip_far = False
    variants["no clipping"] = run()
    m.render_clip_far = True
    m.render_samples_per_round = 8   # fixed n_step slots per ray, the reference's layout (no compaction)
    variants["slot layout"] = run()
    m.render_samples_per_round = 64
    for what, (a, b) in variants.items():
        for k in ("image", "depth", "t"):
            assert torch.equal(base[0][k], a[k]), (name, what, k)
        for k in ("image", "weights", "weights_edit", "depth", "depth_edit"):
            assert torch.equal(base[1][k], b[k]), (name, what, "distill", k)
    assert variants["slot layout"][0]["num_points"] > base[0]["num_points"]   # the slots that compaction does not march
    # the box: every occupied cell (all cascades) inside, in world coordinates
    box = m.occupied_box().cpu().numpy()
    H, C_ = m.grid_size, m.cascade
    bits = np.unpackbits(m.density_bitfield.cpu().numpy(), bitorder="little").reshape(C_, H ** 3)
    coords = raymarching.morton3D_invert(torch.arange(H ** 3, dtype=torch.int32, device=dev)).cpu().numpy()
    for c in range(C_):
        occ = coords[bits[c] == 1]
        if len(occ) == 0:
            continue
        half = min(2 ** c, sc.bound)
        cell = 2.0 * half / H
        lo, hi = -half + occ.min(0) * cell, -half + (occ.max(0) + 1) * cell
        assert (box[:3] <= lo - 1.9 * cell).all() and (box[3:] >= hi + 1.9 * cell).all(), (c, box, lo, hi)


def test_auto_render_schedule_edge_cases(dev):
    """"auto" against the reference schedule where the fast path has to cope or step aside: one ray, rays that all miss the
    occupied box, an empty occupancy grid, a max_steps cap that cuts rays off (auto must land on the reference schedule), dt_gamma > 0
    with several cascades (non-constant step: no closed-form windows), and a ray count that is not a multiple of anything."""
    from laenerf_b200.nerf import NeRFNetwork

    def model(name, seed=9):
        sc = scene(name)
        torch.manual_seed(seed)
        m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=5.0).to(dev)
        with torch.no_grad():
            m.encoder.embeddings.uniform_(-0.5, 0.5)
        m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
        m.eval()
        return m

    def both(m, ro, rd, **kw):
        outs = {}
        for sched in ("reference", "auto"):
            m.render_schedule = sched
            m._auto_fast_ok = True
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                outs[sched] = m.render(ro, rd, perturb=False, bg_color=1, scale_depth=False, **kw)
        for k in ("image", "depth", "t"):
            assert torch.equal(outs["reference"][k], outs["auto"][k]), (k, kw)
        return outs

    m = model("lego")
    _, ro, rd, _ = scene_rays("lego", 5003, 41)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    both(m, ro[:1], rd[:1])                                   # one ray
    both(m, ro, rd)                                           # 5003 rays
    o = both(m, ro, rd, max_steps=16)                         # the cap cuts rays off
    assert o["auto"]["schedule"] == "reference"
    away = both(m, ro, -rd)                                   # every ray points away from the object: all miss the occupied box
    assert float(away["auto"]["t"].abs().max()) == 0.0 and torch.equal(away["auto"]["image"], torch.ones_like(away["auto"]["image"]))
    m.density_bitfield.zero_()                                # nothing occupied at all: the box is (+inf, -inf)
    empty = both(m, ro, rd)
    assert float(empty["auto"]["t"].abs().max()) == 0.0
    mb = model("bonsai")
    _, ro, rd, _ = scene_rays("bonsai", 20011, 43)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    both(mb, ro, rd, dt_gamma=1.0 / 128)                      # growing steps, five cascades


@pytest.mark.parametrize("name", ["lego", "flower", "bonsai"])
def test_training_march_clipped_to_the_occupied_box_is_identical(dev, name):
    """lnrf_march_rays_train_clipped against lnrf_march_rays_train: positions, directions, deltas, ray table and counter bit for bit
    (rays end where they leave the box around the occupied cells; nothing is ever sampled beyond it), with and without perturbation,
    dt_gamma = 0 and > 0."""
    from laenerf_b200 import raymarching
    from laenerf_b200.nerf import NeRFNetwork
    sc = scene(name)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    _, ro, rd, _ = scene_rays(name, 20000, 17)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, m.aabb_train, m.min_near)
    box = m.occupied_box()
    assert torch.isfinite(box).all() and bool((box[3:] > box[:3]).all())
    for perturb, dt_gamma in ((False, 0.0), (True, 0.0), (True, 1.0 / 256)):
        outs = []
        for b in (None, box):
            torch.manual_seed(123)
            counter = torch.zeros(2, dtype=torch.int32, device=dev)
            o = raymarching.march_rays_train(ro, rd, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter, -1, perturb, 128,
                                             True, dt_gamma, 1024, None, b)
            outs.append((*o, counter))
        for x, y in zip(*outs):
            assert torch.equal(x, y), (name, perturb, dt_gamma)
        assert int(outs[0][-1][0]) > 1000
