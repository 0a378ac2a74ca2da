import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from laenerf_b200 import _native as N, raymarching
from laenerf_b200.nerf import NeRFNetwork
dev = torch.device("cuda", 0)
torch.manual_seed(7)
a = NeRFNetwork(bound=1, density_thresh=0.01).to(dev)
with torch.no_grad():
    a.encoder.embeddings.uniform_(-0.5, 0.5)
H = 128
n = H ** 3
u = torch.rand(n, 3, device=dev)
xyzs = torch.empty(n, 3, device=dev); idx = torch.empty(n, dtype=torch.int32, device=dev)
N.check(N.lib().lnrf_occupancy_points(None, N.ptr(u), n, H, 1.0, N.ptr(xyzs), N.ptr(idx), N.stream()))
ar = torch.arange(H, dtype=torch.int32, device=dev)
xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
ind = raymarching.morton3D(coords)
x0 = 2 * coords.float() / (H - 1) - 1
hgs = 1 / H
ref = x0 * (1 - hgs)
ref += (u * 2 - 1) * hgs
print("idx equal", torch.equal(ind, idx), "xyz max diff", float((ref - xyzs).abs().max()), "n differing", int((ref != xyzs).sum()))
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    s1 = a._density_scaled(xyzs)
    a.fused = False
    s2 = a._density_scaled(xyzs)
print("sigma max abs diff", float((s1 - s2).abs().max()), "max rel", float(((s1 - s2).abs() / s2.abs()).max()), s1[:4], s2[:4])
