#!/usr/bin/env python
"""Render time of rank 0's tile share of the lego-shape frame for world sizes 1, 2, 4, 8 -- on ONE GPU (predicts the render scaling
without an 8-GPU box: the ranks are independent up to the final gather).  Diagnostic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.parallel import tile_shard_indices
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
for name in ("lego", "bonsai"):
    sc = make_scene(name, seed=0, n_poses=2)
    torch.manual_seed(0)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    m.eval()
    ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
    for world in (1, 2, 4, 8):
        mine = tile_shard_indices(sc.H, sc.W, 0, world).numpy() if world > 1 else np.arange(sc.H * sc.W)
        o, d = torch.from_numpy(ro[mine]).to(dev), torch.from_numpy(rd[mine]).to(dev)
        for sched in ("auto", "fast", "reference"):
            m.render_schedule, m._auto_fast_ok = sched, True
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                for _ in range(2):
                    out = m.render(o, d, perturb=False, bg_color=1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(3):
                    out = m.render(o, d, perturb=False, bg_color=1)
                e1.record(); torch.cuda.synchronize()
            print("%-6s world %d  %-9s %7.2f ms  rounds %3d  slots %9d  (%s)" % (name, world, sched, e0.elapsed_time(e1) / 3, out["rounds"], out["num_points"], out.get("schedule")))
