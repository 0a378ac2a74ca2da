"""Diagnostic: where does the fix-up pass of the "auto" render schedule differ from the reference schedule?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cases import scene, scene_rays
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200 import raymarching
dev = torch.device("cuda")
name, n = "bonsai", 40000
sc = scene(name)
torch.manual_seed(3)
mb = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=5.0).to(dev)
with torch.no_grad():
    mb.encoder.embeddings.uniform_(-0.5, 0.5)
mb.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
_, ro, rd, _ = scene_rays(name, n, 29)
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
mb.eval()
nears, fars = raymarching.near_far_from_aabb(ro, rd, mb.aabb_infer, mb.min_near)
args = (ro, rd, nears, fars, mb.density_bitfield, None, 0, False, 1024, 1e-4, 8)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    r = mb._render_rounds_on("reference", *args, track=True)
    f = mb._render_rounds_on("fast", *args, track=True)
    mb.render_schedule = "auto"
    a = mb._render_rounds_device(*args)
print("rounds ref/fast", r["rounds"], f["rounds"], "inexact rays ref/fast", r["inexact_rays"], f["inexact_rays"], "cap", r["cap_cut"], f["cap_cut"])
print(a["schedule"])
fr, ff = r["ray_flags"].bool(), f["ray_flags"].bool()
print("flag sets: ref", int(fr.sum()), "fast", int(ff.sum()), "ref&~fast", int((fr & ~ff).sum()), "fast&~ref", int((ff & ~fr).sum()))
sr, sf = r["ray_steps"], f["ray_steps"]
ds = sr != sf
print("death samples differ:", int(ds.sum()), " among unflagged(fast):", int((ds & ~ff).sum()), " among unflagged(either):", int((ds & ~ff & ~fr).sum()))
for k in ("image", "depth", "weights_sum"):
    x, y, z = r[k], f[k], a[k]
    d_rf = (x != y).reshape(n, -1).any(1)
    d_ra = (x != z).reshape(n, -1).any(1)
    print(k, "ref!=fast", int(d_rf.sum()), "of which unflagged", int((d_rf & ~ff).sum()), "| ref!=auto", int(d_ra.sum()), "of which unflagged", int((d_ra & ~ff).sum()),
          "maxabs", float((x - z).abs().max()))
hist = torch.bincount(sr.long().clamp(max=1096), minlength=1097).cpu().numpy()
seq_ref = NeRFNetwork._reference_sequence(hist, n, 1024)
print("seq from the reference run's death samples:", len(seq_ref), "rounds; reference run took", r["rounds"])
hist_f = torch.bincount(sf.long().clamp(max=1096), minlength=1097).cpu().numpy()
seq_f = NeRFNetwork._reference_sequence(hist_f, n, 1024)
same = sum(1 for x, y in zip(seq_ref, seq_f) if x == y)
print("seq from fast:", len(seq_f), "first mismatch at", next((i for i, (x, y) in enumerate(zip(seq_ref, seq_f)) if x != y), None))
bad = (r["image"] != a["image"]).any(1).nonzero().flatten()[:8]
for i in bad.tolist():
    print(i, "flag", int(ff[i]), int(fr[i]), "steps ref/fast", int(sr[i]), int(sf[i]), r["image"][i].tolist(), a["image"][i].tolist())

import time
def timed(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for name_, n_ in (("bonsai", 640000), ("flower", 190512)):
        sc = scene(name_)
        torch.manual_seed(3)
        mb = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_scale=5.0).to(dev)
        mb.encoder.embeddings.uniform_(-0.5, 0.5)
        mb.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
        _, ro, rd, _ = scene_rays(name_, n_, 29)
        ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
        mb.eval()
        nears, fars = raymarching.near_far_from_aabb(ro, rd, mb.aabb_infer, mb.min_near)
        args = (ro, rd, nears, fars, mb.density_bitfield, None, 0, False, 1024, 1e-4, 8)
        t_fast = timed(lambda: mb._render_rounds_on("fast", *args, track=True))
        mb.render_schedule = "auto"
        t_auto = timed(lambda: mb._render_rounds_device(*args))
        os.environ["LNRF_FIXUP_ROUNDS"] = "1"
        t_auto_r = timed(lambda: mb._render_rounds_device(*args))
        del os.environ["LNRF_FIXUP_ROUNDS"]
        t_ref = timed(lambda: mb._render_rounds_on("reference", *args), 1)
        a = mb._render_rounds_device(*args)
        print(name_, n_, "rays: fast %.2f ms | auto (one-pass fix-up) %.2f ms | auto (fix-up by rounds) %.2f ms | reference %.2f ms |" % (t_fast, t_auto, t_auto_r, t_ref), a["schedule"][:60], "slots", a["slots"])
