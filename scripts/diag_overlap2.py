#!/usr/bin/env python
"""Which kernels of the training step can hide the NEXT batch's occupancy march (latency-bound, parameter-independent)?
Each candidate alone, the march chain alone, and both on two streams -- every variant captured in a CUDA graph (diagnostic, GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from laenerf_b200 import _native as N
from laenerf_b200.gridencoder import _offsets_host
from laenerf_b200.nerf import NeRFNetwork, TrainStep
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
lib = N.lib()
sc = make_scene("lego", seed=0, n_poses=2)
m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W, N=4096, rng=np.random.default_rng(0))
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
gt = torch.rand(4096, 3, device=dev)
step = TrainStep(m)
for _ in range(3):
    step(ro, rd, gt)
m.update_mean_count()
for _ in range(3):
    step(ro, rd, gt)
mr = m.march_train(ro, rd, perturb=True)
into = {k: v.clone() for k, v in mr.items()}
xyzs, dirs = mr["xyzs"], mr["dirs"]
M = xyzs.shape[0]
print("rows", M, "live", int(mr["counter"][0]))
ns, nc, ds = 2, 3, 1.0
enc = m.encoder
emb16 = enc.embeddings.detach().half()
off_h = _offsets_host(enc.offsets)
L, S, H = 16, float(np.log2(enc.per_level_scale)), 16
feat = torch.empty(M, 32, dtype=torch.half, device=dev)
ws, wc = m.sigma_net.weights.detach().half(), m.color_net.weights.detach().half()
sig, rgb, h = torch.empty(M, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 16, dtype=torch.half, device=dev)
gsig, grgb = torch.randn(M, device=dev) * 1e-2, torch.randn(M, 3, device=dev) * 1e-1
nbytes = lib.lnrf_nerf_wgrad_scratch_bytes(ns, nc)
scratch = torch.empty(nbytes // 4, device=dev)
genc, gws, gwc = torch.empty_like(feat), torch.empty_like(ws), torch.empty_like(wc)
gemb = torch.zeros_like(emb16)
cnt = mr["counter"]

def s(): return torch.cuda.current_stream().cuda_stream
def enc_fwd():
    N.check(lib.lnrf_grid_encode_forward_world(N.ptr(xyzs), float(sc.bound), N.ptr(emb16), N.ptr(off_h), N.ptr(feat), M, N.ptr(cnt), L, S, H, 0, 0, 0, N.F16, s()))
def nerf_fwd():
    N.check(lib.lnrf_nerf_forward_lean(N.ptr(feat), N.ptr(dirs), N.ptr(ws), N.ptr(wc), M, N.ptr(cnt), ns, nc, ds, N.ptr(h), N.ptr(sig), N.ptr(rgb), s()))
def nerf_bwd():
    N.check(lib.lnrf_nerf_backward_recompute(N.ptr(gsig), N.ptr(grgb), N.ptr(rgb), N.ptr(h), N.ptr(feat), N.ptr(dirs), N.ptr(ws), N.ptr(wc), M, N.ptr(cnt), ns, nc, ds,
                                             N.ptr(genc), N.ptr(gws), N.ptr(gwc), 0, N.ptr(scratch), nbytes, s()))
def enc_bwd():
    N.check(lib.lnrf_grid_encode_backward_world(N.ptr(genc), N.ptr(xyzs), float(sc.bound), N.ptr(off_h), N.ptr(gemb), M, N.ptr(cnt), L, S, H, 0, 0, 0, N.F16, s()))
def adam():
    step.optimizer.step()
def march():
    step.march(ro, rd, True, into=into)
enc_fwd(); nerf_fwd(); torch.cuda.synchronize()
side = torch.cuda.Stream()

def graph_time(main_fns, side_fns, reps=20):
    g = torch.cuda.CUDAGraph()
    warm = torch.cuda.Stream()
    warm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(warm):
        for f in main_fns + side_fns:
            f()
    torch.cuda.current_stream().wait_stream(warm)
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for _ in range(reps):
            if side_fns:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for f in side_fns:
                        f()
            for f in main_fns:
                f()
            if side_fns:
                main.wait_stream(side)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3

t_m = graph_time([march], [])
print("march chain alone %.1f us" % t_m)
for name, fns in (("enc_bwd", [enc_bwd]), ("adam", [adam]), ("enc_fwd", [enc_fwd]), ("nerf_fwd", [nerf_fwd]), ("nerf_bwd", [nerf_bwd]),
                  ("enc_bwd+adam", [enc_bwd, adam]), ("enc_fwd+nerf_fwd", [enc_fwd, nerf_fwd])):
    a = graph_time(fns, [])
    b = graph_time(fns, [march])
    print("%-18s alone %6.1f us | beside the march %6.1f us | hidden %5.1f of %.1f us" % (name, a, b, a + t_m - b, t_m))
