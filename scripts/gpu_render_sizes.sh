#!/bin/bash
# frame time of the device-driven render loop for one rank's share of a tile-sharded frame (1 GPU), plus its launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/render_sizes.log
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200.parallel import tile_shard_indices
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
sc = make_scene("lego", seed=0, n_poses=4)
torch.manual_seed(0)
model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
model.eval()
model.render_schedule = os.environ.get('LNRF_RENDER_SCHEDULE', 'fast')
model.render_samples_per_round = int(os.environ.get('LNRF_RENDER_SPR', '32'))
for world in (1, 2, 4, 8):
    mine = tile_shard_indices(sc.H, sc.W, 0, world).numpy() if world > 1 else np.arange(sc.H * sc.W)
    o, d = torch.from_numpy(ro[mine]).to(dev), torch.from_numpy(rd[mine]).to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(2): out = model.render(o, d, perturb=False, bg_color=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        for _ in range(3): out = model.render(o, d, perturb=False, bg_color=1)
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"share 1/{world}: rays {len(mine):7d}  ms/frame {e0.elapsed_time(e1)/3:8.3f} (wall {(t1-t0)/3*1e3:8.3f})  rounds {out.get('rounds')}  slots {out['num_points']}")
PY

