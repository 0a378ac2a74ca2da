#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_render.csv python scripts/profile_step.py --steps 0 --render-rays 640000 > gpurun_out/ncu_render.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_render.log
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np, time
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.scene import get_rays_np, make_scene
from laenerf_b200 import raymarching
dev = torch.device("cuda", 0)
sc = make_scene("lego", seed=0, n_poses=4)
torch.manual_seed(0)
model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
ro, rd, _ = get_rays_np(sc.poses[1], sc.intrinsics, sc.H, sc.W)
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
model.eval()
# instrument the loop: per-round n_alive, n_step, real samples
model.device_loop = False
rounds = []
orig = raymarching.march_rays
def spy(n_alive, n_step, *a, **k):
    out = orig(n_alive, n_step, *a, **k)
    rounds.append((n_alive, n_step, out[0].shape[0], int((out[2][:, 0] > 0).sum().item())))
    return out
raymarching.march_rays = spy
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    model.render(ro, rd, perturb=False)
raymarching.march_rays = orig
print("rounds", len(rounds), "slots", sum(r[2] for r in rounds), "real samples", sum(r[3] for r in rounds))
for r in rounds[:12] + rounds[-4:]: print(r)
model.device_loop = True
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(2): model.render(ro, rd, perturb=False)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(3): model.render(ro, rd, perturb=False)
    torch.cuda.synchronize(); print("ms/frame", (time.time()-t)/3*1e3)
PY
