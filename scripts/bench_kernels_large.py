#!/usr/bin/env python
"""march / encode / composite at the sizes where an HBM roofline is meaningful (the 4096-ray training
launches move 6-130 MB each and are latency-bound; "the 640 000-ray render rounds and M >= 10^6 encode/composite launches
are the sizes at which the >= 60 % HBM target is meaningful").  One lego-shape batch of 65 536 training rays (~3.8 M samples)
through the drop-in ops, CUDA events, algorithmic bytes of SURVEY 8d against MEASURED_PEAKS.json.
    python scripts/bench_kernels_large.py > profiles/<tag>_large_batch.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from laenerf_b200 import raymarching
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.scene import get_rays_np, make_scene

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = float(peaks["hbm_gbs"])
N = 65536


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


sc = make_scene("lego", seed=0, n_poses=8)
torch.manual_seed(0)
model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
rng = np.random.default_rng(0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W, N=N, rng=rng)
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
nears, fars = raymarching.near_far_from_aabb(ro, rd, model.aabb_train, model.min_near)
counter = torch.zeros(2, dtype=torch.int32, device=dev)
# size the sample buffer like the reference does after the first steps: mean_count = the realised total
xyzs, dirs, deltas, rays = raymarching.march_rays_train(ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars,
                                                        counter, -1, True, 128, True, 0, 1024)
M_real = int(counter[0].item())
mean_count = M_real


def march():
    counter.zero_()
    return raymarching.march_rays_train(ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars, counter,
                                        mean_count, True, 128, False, 0, 1024)


xyzs, dirs, deltas, rays = march()
M = int(xyzs.shape[0])
out = {"gpu": torch.cuda.get_device_name(0), "rays": N, "samples": M_real, "rows": M, "hbm_peak_gbs": HBM, "kernels": {}}


def rec(name, ms, byts):
    out["kernels"][name] = dict(ms=ms, algorithmic_bytes=byts, achieved_gbs=byts / ms / 1e6, frac_of_hbm_peak=byts / ms / 1e6 / HBM)


rec("march_rays_train", timed(march), 48 * N + 32 * M_real)

enc = model.encoder
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    feat = enc(xyzs, bound=model.bound)
    rec("grid_encode_forward (fp16 table)", timed(lambda: enc(xyzs, bound=model.bound)), 588 * M)
g = torch.randn(M, 32, device=dev).half() * 1e-3
with torch.autocast("cuda", dtype=torch.float16):
    f2 = enc(xyzs, bound=model.bound)


def enc_bwd():
    enc.embeddings.grad = None
    f2.backward(g, retain_graph=True)


rec("grid_encode_backward (fp16 grads, incl. the 24.5 MB clear torch does)", timed(enc_bwd), 588 * M)

sig = torch.rand(M, device=dev) * 20
rgb = torch.rand(M, 3, device=dev)
rec("composite_rays_train forward", timed(lambda: raymarching.composite_rays_train(sig, rgb, deltas, rays, 1e-4)), 32 * N + 24 * M_real)
sg, rg = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
ws, dp, img = raymarching.composite_rays_train(sg, rg, deltas, rays, 1e-4)
gi, gw = torch.randn_like(img), torch.randn_like(ws)


def comp_bwd():
    sg.grad = rg.grad = None
    torch.autograd.backward([img, ws], [gi, gw], retain_graph=True)


rec("composite_rays_train backward (incl. the two zeros_like the wrapper keeps)", timed(comp_bwd), 44 * N + 40 * M_real)
gt = torch.rand(N, 3, device=dev)
rec("composite + blend + MSE forward (row f-5)", timed(lambda: raymarching.composite_loss_train(sig, rgb, deltas, rays, gt, 1, nears, fars, 1e-4)),
    64 * N + 24 * M_real)
print(json.dumps(out, indent=1))
