#!/usr/bin/env python
"""Can the MLP backward (TMEM-read-bound, 1 CTA/SM) and the hash-grid backward (L2-atomic / LSU-bound) share the SMs?
Times both alone, launched together on two streams, and as a chunked two-stream pipeline (diagnostic, GPU)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from laenerf_b200 import _native as N
from laenerf_b200.gridencoder import GridEncoder, _offsets_host
from laenerf_b200.nerf import NeRFNetwork
from laenerf_b200 import raymarching
from laenerf_b200.scene import get_rays_np, make_scene
dev = torch.device("cuda", 0)
lib = N.lib()
sc = make_scene("lego", seed=0, n_poses=2)
m = NeRFNetwork(bound=sc.bound, min_near=sc.min_near).to(dev)
m.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W, N=4096, rng=np.random.default_rng(0))
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
m.train()
mr = m.march_train(ro, rd, perturb=True)
xyzs, dirs = mr["xyzs"], mr["dirs"]
M = xyzs.shape[0] // 512 * 512
xyzs, dirs = xyzs[:M].contiguous(), dirs[:M].contiguous()
print("rows", M)
ns, nc, ds = 2, 3, 1.0
enc = m.encoder
emb16 = enc.embeddings.detach().half()
off_h = _offsets_host(enc.offsets)
L, S, H = 16, float(np.log2(enc.per_level_scale)), 16
feat = torch.empty(M, 32, dtype=torch.half, device=dev)
N.check(lib.lnrf_grid_encode_forward_world(N.ptr(xyzs), float(sc.bound), N.ptr(emb16), N.ptr(off_h), N.ptr(feat), M, None, L, S, H, 0, 0, 0, N.F16, None))
ws, wc = m.sigma_net.weights.detach().half(), m.color_net.weights.detach().half()
sig, rgb, h = torch.empty(M, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 16, dtype=torch.half, device=dev)
gsig, grgb = torch.randn(M, device=dev) * 1e-2, torch.randn(M, 3, device=dev) * 1e-1
nbytes = lib.lnrf_nerf_wgrad_scratch_bytes(ns, nc)
scratch = torch.empty(nbytes // 4, device=dev)
genc, gws, gwc = torch.empty_like(feat), torch.empty_like(ws), torch.empty_like(wc)
gemb = torch.zeros_like(emb16)
A, B = torch.cuda.Stream(), torch.cuda.Stream()

def nerf_fwd(lo, n, st):
    N.check(lib.lnrf_nerf_forward_lean(feat[lo:].data_ptr(), dirs[lo:].data_ptr(), N.ptr(ws), N.ptr(wc), n, None, ns, nc, ds, h[lo:].data_ptr(),
                                       sig[lo:].data_ptr(), rgb[lo:].data_ptr(), st.cuda_stream))
def enc_fwd(lo, n, st):
    N.check(lib.lnrf_grid_encode_forward_world(xyzs[lo:].data_ptr(), float(sc.bound), N.ptr(emb16), N.ptr(off_h), feat[lo:].data_ptr(), n, None, L, S, H,
                                               0, 0, 0, N.F16, st.cuda_stream))
def nerf_bwd(lo, n, st):
    N.check(lib.lnrf_nerf_backward_recompute(gsig[lo:].data_ptr(), grgb[lo:].data_ptr(), rgb[lo:].data_ptr(), h[lo:].data_ptr(), feat[lo:].data_ptr(),
                                             dirs[lo:].data_ptr(), N.ptr(ws), N.ptr(wc), n, None, ns, nc, ds, genc[lo:].data_ptr(), N.ptr(gws), N.ptr(gwc), 0,
                                             N.ptr(scratch), nbytes, st.cuda_stream))
def enc_bwd(lo, n, st):
    N.check(lib.lnrf_grid_encode_backward_world(genc[lo:].data_ptr(), xyzs[lo:].data_ptr(), float(sc.bound), N.ptr(off_h), N.ptr(gemb), n, None, L, S, H,
                                                0, 0, 0, N.F16, st.cuda_stream))
nerf_fwd(0, M, A); torch.cuda.synchronize()

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

cur = torch.cuda.current_stream()
def join():
    cur.wait_stream(A); cur.wait_stream(B)
def fork():
    A.wait_stream(cur); B.wait_stream(cur)

def both(f1, f2):
    def run():
        fork(); f1(0, M, A); f2(0, M, B); join()
    return run
def alone(f, st):
    def run():
        fork(); f(0, M, st); join()
    return run
def pipeline(f1, f2, chunks):
    step = M // chunks // 128 * 128
    def run():
        fork()
        lo = 0
        for c in range(chunks):
            n = step if c < chunks - 1 else M - lo
            f1(lo, n, A)
            ev = torch.cuda.Event(); ev.record(A); B.wait_event(ev)
            f2(lo, n, B)
            lo += n
        join()
    return run

print("backward:  nerf alone %.1f us | enc alone %.1f us | together (independent) %.1f us" %
      (timed(alone(nerf_bwd, A)), timed(alone(enc_bwd, B)), timed(both(nerf_bwd, enc_bwd))))
for c in (2, 3, 4, 6, 8):
    print("  pipeline nerf_bwd -> enc_bwd, %d chunks: %.1f us" % (c, timed(pipeline(nerf_bwd, enc_bwd, c))))
print("forward:   enc alone %.1f us | nerf alone %.1f us | together (independent) %.1f us" %
      (timed(alone(enc_fwd, A)), timed(alone(nerf_fwd, B)), timed(both(enc_fwd, nerf_fwd))))
for c in (2, 3, 4, 6, 8):
    print("  pipeline enc_fwd -> nerf_fwd, %d chunks: %.1f us" % (c, timed(pipeline(enc_fwd, nerf_fwd, c))))
