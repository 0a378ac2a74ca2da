#!/bin/bash
# frame time of the device-driven render loop versus the number of rays (what one rank of a tile-sharded frame sees)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/render_sizes.log
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
sc = make_scene("lego", seed=0, n_poses=4)
torch.manual_seed(0)
model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
model.eval()
for n in (640000, 320000, 160000, 80000, 8192):
    o, d = torch.from_numpy(ro[:n]).to(dev), torch.from_numpy(rd[:n]).to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(2): out = model.render(o, d, perturb=False, bg_color=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3): out = model.render(o, d, perturb=False, bg_color=1)
        e1.record(); torch.cuda.synchronize()
    print(f"rays {n:7d}  ms/frame {e0.elapsed_time(e1)/3:8.3f}  rounds {out.get('rounds')}  slots {out['num_points']}")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_render80k.csv python scripts/profile_step.py --steps 0 --render-rays 80000 > gpurun_out/ncu_render80k.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_render80k.log
