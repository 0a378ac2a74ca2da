import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_fused import _update_extra_state_torch
from laenerf_b200.nerf import NeRFNetwork
dev = torch.device("cuda", 0)
torch.manual_seed(7)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
a = NeRFNetwork(bound=B, density_thresh=0.01).to(dev)
with torch.no_grad():
    a.encoder.embeddings.uniform_(-0.5, 0.5)
b = NeRFNetwork(bound=B, density_thresh=0.01).to(dev)
b.load_state_dict(a.state_dict()); b.fused = False
for it in range(2):
    with torch.autocast("cuda", dtype=torch.float16):
        torch.manual_seed(100 + it); a.update_extra_state()
        torch.manual_seed(100 + it); _update_extra_state_torch(b)
    d = (a.density_grid - b.density_grid).abs()
    print(it, "max diff", float(d.max()), "n>1e-4", int((d > 1e-4).sum()), "mean", a.mean_density, b.mean_density,
          "grid stats", float(a.density_grid.min()), float(a.density_grid.max()), float(b.density_grid.min()), float(b.density_grid.max()))
    j = int(d.view(-1).argmax()); print("  at", j, float(a.density_grid.view(-1)[j]), float(b.density_grid.view(-1)[j]))
    for c in range(a.cascade): print("   cascade", c, "n differing", int((a.density_grid[c] != b.density_grid[c]).sum()))
