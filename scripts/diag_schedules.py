#!/usr/bin/env python
"""Are the two round schedules of the device render loop bit-identical?  (diagnostic, GPU)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
for name, ds in (("lego", 1.0), ("lego", 30.0), ("bonsai", 1.0), ("flower", 10.0)):
    sc = make_scene(name, seed=0, n_poses=2)
    torch.manual_seed(0)
    m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=ds).to(dev)
    with torch.no_grad():
        m.encoder.embeddings.uniform_(-0.5, 0.5)
    m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    m.eval()
    outs = {}
    for sched in ("reference", "fast"):
        m.render_schedule = sched
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            o = m.render(ro, rd, perturb=False, bg_color=1, T_thresh=1e-4, scale_depth=False)
        outs[sched] = o
    a, b = outs["reference"], outs["fast"]
    for k in ("image", "depth", "t"):
        x, y = a[k], b[k]
        diff = (x != y) & ~(torch.isnan(x) & torch.isnan(y))
        nd = int(diff.sum())
        print(name, ds, k, "rounds", a["rounds"], b["rounds"], "differing elements:", nd, "of", x.numel(), "max abs diff", float((x - y).abs().nan_to_num().max()))
        if nd and k == "t":
            idx = diff.nonzero().flatten()[:5].tolist()
            for i in idx:
                print("   ray", i, float(x[i]), float(y[i]), "depth", float(a["depth"][i]), float(b["depth"][i]))
