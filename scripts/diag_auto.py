import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200 import raymarching
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
sc = make_scene("lego", seed=0, n_poses=8)
torch.manual_seed(0)
m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
m.eval()
for pose in range(4):
    ro, rd, _ = get_rays_np(sc.poses[pose], sc.intrinsics, sc.H, sc.W)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, m.aabb_infer, m.min_near)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        t = m._render_rounds_on("fast", ro, rd, nears, fars, m.density_bitfield, None, 0, False, 1024, 1e-4)
        r = m._render_rounds_on("reference", ro, rd, nears, fars, m.density_bitfield, None, 0, False, 1024, 1e-4)
    ok = nears < 1e30
    print("pose", pose, "flag", t["schedule_dependent"], "rays flagged", t["inexact_rays"], "rounds", t["rounds"], "| reference-schedule flagged", r["inexact_rays"],
          "| near min %.3f far max %.3f" % (float(nears[ok].min()), float(fars[ok].max())), "identical image", bool(torch.equal(t["image"], r["image"])),
          "rays_t equal", bool(torch.equal(t["rays_t"], r["rays_t"])))
