#!/usr/bin/env python
"""Per-step cycle breakdown of k_nerf_bwd (block 0, tile set 0): waiting for the MMA batch vs epilogue work (diagnostic, GPU)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from laenerf_b200 import _native as N
dev = torch.device("cuda", 0)
lib = N.lib()
raw = C.CDLL(N.SO_PATH)
M, ns, nc, ds = 128 * 1977, 2, 3, 1.0
g = torch.Generator(device=dev).manual_seed(0)
enc = (torch.randn(M, 32, device=dev, generator=g) * 0.5).half()
dirs = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1)
ws = ((torch.rand(64 * (32 + 64 * (ns - 1) + 16), device=dev, generator=g) - 0.5) * 0.5).half()
wc = ((torch.rand(64 * (32 + 64 * (nc - 1) + 16), device=dev, generator=g) - 0.5) * 0.5).half()
gsig = torch.randn(M, device=dev, generator=g) * 1e-2
grgb = torch.randn(M, 3, device=dev, generator=g) * 1e-1
sig, rgb, h = torch.empty(M, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 16, dtype=torch.half, device=dev)
N.check(lib.lnrf_nerf_forward_lean(N.ptr(enc), N.ptr(dirs), N.ptr(ws), N.ptr(wc), M, None, ns, nc, ds, N.ptr(h), N.ptr(sig), N.ptr(rgb), None))
nbytes = lib.lnrf_nerf_wgrad_scratch_bytes(ns, nc)
scratch = torch.empty(nbytes // 4, device=dev)
genc, gws, gwc = torch.empty_like(enc), torch.empty_like(ws), torch.empty_like(wc)
def run():
    N.check(lib.lnrf_nerf_backward_recompute(N.ptr(gsig), N.ptr(grgb), N.ptr(rgb), N.ptr(h), N.ptr(enc), N.ptr(dirs), N.ptr(ws), N.ptr(wc), M, None,
                                             ns, nc, ds, N.ptr(genc), N.ptr(gws), N.ptr(gwc), 0, N.ptr(scratch), nbytes, None))
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print("k_nerf_bwd + reduce: %.1f us per call at M = %d" % (e0.elapsed_time(e1) * 100, M))
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
raw.lnrf_debug_set_nerf_bwd_counters.argtypes = [C.c_void_p]
raw.lnrf_debug_set_nerf_bwd_counters(dbg.data_ptr())
run(); torch.cuda.synchronize()
raw.lnrf_debug_set_nerf_bwd_counters(None)
d = dbg.tolist()
tiles = (1977 + 295) // 296   # tiles of block 0, set 0
nsteps = 2 * (ns + nc) + 2
names = ["C fwd0", "C fwd1", "C fwd2", "C bwd3", "C bwd2", "C bwd1", "C in", "S fwd0", "S fwd1", "S bwd2", "S bwd1", "S in"]
tot = 0
for s in range(nsteps):
    w, e = d[2 * s] / tiles, d[2 * s + 1] / tiles
    tot += w + e
    print("step %2d %-7s wait %6.0f cyc   epilogue %6.0f cyc" % (s, names[s] if s < len(names) else "", w, e))
print("per tile %.0f cycles (%d tiles)" % (tot, tiles))
