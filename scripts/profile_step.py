#!/usr/bin/env python
"""Workload for ncu: warm up the lego-shape training step, then run a few steps (and optionally render rounds)
between cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures steady-state launches only."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from laenerf_b200.nerf import NeRFNetwork, TrainStep
from laenerf_b200.scene import get_rays_np, make_scene

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=6)
ap.add_argument("--render-rays", type=int, default=0, help="also profile one render of this many rays")
ap.add_argument("--render-shard", type=int, default=0, help="render rank 0's share of a frame tile-sharded over this many ranks")
ap.add_argument("--render-schedule", default="fast", choices=["fast", "reference", "auto"])
ap.add_argument("--scene", default="lego", choices=["lego", "flower", "bonsai"])
args = ap.parse_args()

dev = torch.device("cuda", 0)
sc = make_scene(args.scene, seed=0, n_poses=4)
torch.manual_seed(0)
model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
step = TrainStep(model)
rng = np.random.default_rng(0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W, N=4096, rng=rng)
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
gt = torch.rand(4096, 3, device=dev)
for i in range(args.warmup):
    step(ro, rd, gt)
    if i == 0:
        model.update_mean_count()
model.update_mean_count()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(args.steps):
    step(ro, rd, gt)
if args.render_rays > 0 or args.render_shard > 0:
    fo, fd, _ = get_rays_np(sc.poses[1], sc.intrinsics, sc.H, sc.W)
    if args.render_shard > 0:
        from laenerf_b200.parallel import tile_shard_indices
        mine = tile_shard_indices(sc.H, sc.W, 0, args.render_shard).numpy()
        fo, fd = torch.from_numpy(fo[mine]).to(dev), torch.from_numpy(fd[mine]).to(dev)
    else:
        fo, fd = torch.from_numpy(fo[: args.render_rays]).to(dev), torch.from_numpy(fd[: args.render_rays]).to(dev)
    model.eval()
    model.render_schedule = args.render_schedule
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model.render(fo, fd, perturb=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.steps, "steps; mean_count", model.mean_count)
