"""Three interchangeable numpy-in / numpy-out back-ends for the hot-path operators (TEST INFRASTRUCTURE).

    OracleBackend  -- the CPU restatement (oracle/oracle.c through oracle/pyoracle.py); runs anywhere
    OursBackend    -- liblaenerf_b200.so through the raw C ABI (ctypes, torch only for device memory); needs a GPU
    RefBackend     -- the reference's OWN extensions built by oracle/build_ref.py into oracle/_ref/; needs a GPU and
                      is only used by oracle/gen_golden.py to freeze golden vectors (tests never import it)

All three expose the same methods, so a test case is written once (tests/cases.py) and evaluated by any of them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def canonicalize(xyzs, dirs, deltas, rays):
    """Sort a march_rays_train result into ray-id order: returns (counts[N], xyzs, dirs, deltas) where the sample
    arrays are the per-ray segments concatenated in ray-id order (the reference hands out segments in atomic
    arrival order -- SURVEY.md 8a-1 / Appendix A).  Rays that overflowed the buffer must not be present."""
    rays = np.asarray(rays)
    order = np.argsort(rays[:, 0], kind="stable")
    counts = rays[order, 2].astype(np.int64)
    offs = rays[order, 1].astype(np.int64)
    idx = np.concatenate([np.arange(o, o + c) for o, c in zip(offs, counts)]) if counts.sum() > 0 else np.zeros(0, np.int64)
    return counts.astype(np.int32), xyzs[idx], dirs[idx], deltas[idx]


def canonical_rays(counts, base=0):
    counts = np.asarray(counts, np.int64)
    offs = base + np.concatenate([[0], np.cumsum(counts)[:-1]])
    return np.stack([np.arange(len(counts)), offs, counts], -1).astype(np.int32)


# =============================================================================================================
class OracleBackend:
    name = "oracle"

    def __init__(self, device_scales=None):
        from oracle import pyoracle
        self.o = pyoracle
        # per-level scales as the DEVICE evaluates exp2f (golden files carry them); None -> libm exp2f
        self.device_scales = device_scales or {}

    def grid_level_scales(self, L, per_level_scale, H):
        return self.device_scales.get((int(L), int(H)))

    def near_far(self, rays_o, rays_d, aabb, min_near):
        return self.o.near_far_from_aabb(rays_o, rays_d, aabb, min_near)

    def morton3D(self, coords):
        return self.o.morton3D(coords)

    def morton3D_invert(self, idx):
        return self.o.morton3D_invert(idx)

    def packbits(self, grid, thresh):
        return self.o.packbits(grid, thresh)

    def march_train(self, rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter=None):
        return self.o.march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter)

    def composite_train_fwd(self, sigmas, rgbs, deltas, rays, T):
        return self.o.composite_rays_train_forward(sigmas, rgbs, deltas, rays, T)

    def composite_train_bwd(self, gws, gimg, sigmas, rgbs, deltas, rays, ws, image, T):
        return self.o.composite_rays_train_backward(gws, gimg, sigmas, rgbs, deltas, rays, ws, image, T)

    def march(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, M_rows,
              dt_gamma=0.0, max_steps=1024, edit_bitfield=None):
        return self.o.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
                                 M_rows, dt_gamma, max_steps, edit_bitfield)

    def composite(self, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, ws, depth, image, T, wes=None, depth_edit=None,
                  edit_occ=None):
        return self.o.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, ws, depth, image, T, wes,
                                     depth_edit, edit_occ)

    def grid_fwd(self, inputs, emb, offsets, per_level_scale, H, half=False, dy_dx=False, gridtype=0, align=False, interp=0,
                 scales=None):
        emb = np.asarray(emb, np.float32)
        if half:
            emb = emb.astype(np.float16).astype(np.float32)
        return self.o.grid_encode_forward(inputs, emb, offsets, float(np.log2(per_level_scale)), H, dy_dx, gridtype, align, interp,
                                          scales, 1)

    def grid_bwd(self, grad, inputs, offsets, C, per_level_scale, H, half=False, gridtype=0, align=False, interp=0, scales=None):
        grad = np.asarray(grad, np.float32)
        if half:
            grad = grad.astype(np.float16).astype(np.float32)
        return self.o.grid_encode_backward(grad, inputs, offsets, C, float(np.log2(per_level_scale)), H, None, gridtype, align,
                                           interp, scales, 1)

    def ffmlp_fwd(self, inputs, weights, in_dim, out_dim, hidden, n_layers, act=0, out_act=6):
        return self.o.ffmlp_forward(inputs, weights, in_dim, out_dim, hidden, n_layers, act, out_act, True)

    def ffmlp_bwd(self, grad, inputs, weights, fwd_buf, in_dim, out_dim, hidden, n_layers, act=0, calc_grad_inputs=False):
        return self.o.ffmlp_backward(grad, inputs, weights, fwd_buf, in_dim, out_dim, hidden, n_layers, act, calc_grad_inputs)

    def sh(self, dirs, degree):
        return self.o.sh_encode(dirs, degree)


# =============================================================================================================
class _TorchDevice:
    """numpy <-> cuda helpers shared by the two GPU back-ends."""

    def __init__(self):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", 0)

    def t(self, a, dtype=None):
        torch = self.torch
        if a is None:
            return None
        x = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        return x if dtype is None else x.to(dtype)

    def f32(self, a):
        return self.t(np.asarray(a, np.float32))

    def i32(self, a):
        return self.t(np.asarray(a, np.int32))

    def u8(self, a):
        return self.t(np.asarray(a, np.uint8))

    def f16(self, a):
        return self.t(np.asarray(a, np.float32)).half()

    @staticmethod
    def n(x):
        return None if x is None else x.detach().float().cpu().numpy() if x.dtype.is_floating_point else x.detach().cpu().numpy()


class OursBackend(_TorchDevice):
    """liblaenerf_b200.so through the C ABI; every call synchronises and checks for asynchronous CUDA errors."""
    name = "ours"

    def __init__(self):
        super().__init__()
        from laenerf_b200 import _native as N
        self.N = N
        self.lib = N.lib()
        # one scratch per kernel kind, zero before first use (the product wrappers do the same, raymarching._get_scratch)
        self._scratch = self.torch.zeros(1 << 20, dtype=self.torch.int64, device=self.dev)
        self._scratch_compact = self.torch.zeros(1 << 12, dtype=self.torch.int64, device=self.dev)

    def _done(self):
        self.torch.cuda.synchronize()

    def near_far(self, rays_o, rays_d, aabb, min_near):
        N, p = self.N, self.N.ptr
        o, d, a = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3), self.f32(aabb)
        n = o.shape[0]
        nears, fars = self.torch.empty(n, device=self.dev), self.torch.empty(n, device=self.dev)
        N.check(self.lib.lnrf_near_far_from_aabb(p(o), p(d), p(a), n, float(min_near), p(nears), p(fars), None))
        self._done()
        return self.n(nears), self.n(fars)

    def morton3D(self, coords):
        N, p = self.N, self.N.ptr
        c = self.i32(coords).view(-1, 3)
        out = self.torch.empty(c.shape[0], dtype=self.torch.int32, device=self.dev)
        N.check(self.lib.lnrf_morton3D(p(c), c.shape[0], p(out), None))
        self._done()
        return self.n(out)

    def morton3D_invert(self, idx):
        N, p = self.N, self.N.ptr
        c = self.i32(idx).view(-1)
        out = self.torch.empty(c.shape[0], 3, dtype=self.torch.int32, device=self.dev)
        N.check(self.lib.lnrf_morton3D_invert(p(c), c.shape[0], p(out), None))
        self._done()
        return self.n(out)

    def packbits(self, grid, thresh):
        N, p = self.N, self.N.ptr
        g = self.f32(grid).view(-1)
        n = g.shape[0] // 8
        out = self.torch.empty(n, dtype=self.torch.uint8, device=self.dev)
        N.check(self.lib.lnrf_packbits(p(g), n, float(thresh), p(out), None))
        self._done()
        return self.n(out)

    def march_train(self, rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter=None):
        N, p, torch = self.N, self.N.ptr, self.torch
        o, d = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3)
        n = o.shape[0]
        g, ne, fa, no = self.u8(bitfield), self.f32(nears), self.f32(fars), self.f32(noises)
        # poison the outputs: the kernel must overwrite every row (SELF-ZERO contract)
        xyzs = torch.full((M, 3), float("nan"), device=self.dev)
        dirs = torch.full((M, 3), float("nan"), device=self.dev)
        deltas = torch.full((M, 2), float("nan"), device=self.dev)
        rays = torch.full((n, 3), -7, dtype=torch.int32, device=self.dev)
        cnt = self.i32(np.zeros(2) if counter is None else counter)
        nbytes = self.lib.lnrf_march_rays_train_scratch_bytes(n)
        assert nbytes <= self._scratch.numel() * 8
        N.check(self.lib.lnrf_march_rays_train(p(o), p(d), p(g), float(bound), float(dt_gamma), int(max_steps), n, int(C), int(H), M,
                                               p(ne), p(fa), p(xyzs), p(dirs), p(deltas), p(rays), p(cnt), p(no), p(self._scratch),
                                               self._scratch.numel() * 8, None))
        self._done()
        assert int(self._scratch[2:].abs().sum().item()) == 0, "march look-back words not left zeroed"
        return self.n(xyzs), self.n(dirs), self.n(deltas), self.n(rays), self.n(cnt)

    def composite_train_fwd(self, sigmas, rgbs, deltas, rays, T):
        N, p, torch = self.N, self.N.ptr, self.torch
        s, c, dl, r = self.f32(sigmas), self.f32(rgbs), self.f32(deltas), self.i32(rays)
        M, n = s.shape[0], r.shape[0]
        ws, depth, image = (torch.full((n,), float("nan"), device=self.dev), torch.full((n,), float("nan"), device=self.dev),
                            torch.full((n, 3), float("nan"), device=self.dev))
        N.check(self.lib.lnrf_composite_rays_train_forward(p(s), p(c), p(dl), p(r), M, n, float(T), p(ws), p(depth), p(image), None))
        self._done()
        return self.n(ws), self.n(depth), self.n(image)

    def composite_train_bwd(self, gws, gimg, sigmas, rgbs, deltas, rays, ws, image, T, zero_fill=False):
        N, p, torch = self.N, self.N.ptr, self.torch
        s, c, dl, r = self.f32(sigmas), self.f32(rgbs), self.f32(deltas), self.i32(rays)
        M, n = s.shape[0], r.shape[0]
        a, b, w, im = self.f32(gws), self.f32(gimg), self.f32(ws), self.f32(image)
        if zero_fill:
            gs, gc = torch.full((M,), float("nan"), device=self.dev), torch.full((M, 3), float("nan"), device=self.dev)
        else:
            gs, gc = torch.zeros(M, device=self.dev), torch.zeros(M, 3, device=self.dev)
        N.check(self.lib.lnrf_composite_rays_train_backward(p(a), p(b), p(s), p(c), p(dl), p(r), p(w), p(im), M, n, float(T), p(gs),
                                                            p(gc), int(zero_fill), None))
        self._done()
        return self.n(gs), self.n(gc)

    def march(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, M_rows,
              dt_gamma=0.0, max_steps=1024, edit_bitfield=None):
        N, p, torch = self.N, self.N.ptr, self.torch
        o, d = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3)
        ra, rt, g, ne, fa, no = self.i32(rays_alive), self.f32(rays_t), self.u8(bitfield), self.f32(nears), self.f32(fars), self.f32(noises)
        xyzs = torch.full((M_rows, 3), float("nan"), device=self.dev)
        dirs = torch.full((M_rows, 3), float("nan"), device=self.dev)
        deltas = torch.full((M_rows, 2), float("nan"), device=self.dev)
        if edit_bitfield is None:
            N.check(self.lib.lnrf_march_rays(n_alive, n_step, p(ra), p(rt), p(o), p(d), float(bound), float(dt_gamma), int(max_steps),
                                             int(C), int(H), p(g), p(ne), p(fa), p(xyzs), p(dirs), p(deltas), p(no), M_rows, None))
            self._done()
            return self.n(xyzs), self.n(dirs), self.n(deltas)
        eg = self.u8(edit_bitfield)
        occ = torch.full((M_rows,), 3, dtype=torch.uint8, device=self.dev)
        N.check(self.lib.lnrf_march_rays_distill(n_alive, n_step, p(ra), p(rt), p(o), p(d), float(bound), float(dt_gamma),
                                                 int(max_steps), int(C), int(H), p(g), p(eg), p(ne), p(fa), p(xyzs), p(dirs),
                                                 p(deltas), p(occ), p(no), M_rows, None))
        self._done()
        return self.n(xyzs), self.n(dirs), self.n(deltas), self.n(occ)

    def composite(self, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, ws, depth, image, T, wes=None, depth_edit=None,
                  edit_occ=None):
        N, p = self.N, self.N.ptr
        ra, rt = self.i32(rays_alive), self.f32(rays_t)
        s, c, dl = self.f32(sigmas), self.f32(rgbs), self.f32(deltas)
        w, dp, im = self.f32(ws), self.f32(depth), self.f32(image)
        if edit_occ is None:
            N.check(self.lib.lnrf_composite_rays(n_alive, n_step, float(T), p(ra), p(rt), p(s), p(c), p(dl), p(w), p(dp), p(im), None))
            self._done()
            return self.n(ra), self.n(rt), self.n(w), self.n(dp), self.n(im)
        we, de, eo = self.f32(wes), self.f32(depth_edit), self.u8(edit_occ)
        N.check(self.lib.lnrf_composite_rays_distill(n_alive, n_step, float(T), p(ra), p(rt), p(s), p(c), p(dl), p(w), p(we), p(dp),
                                                     p(de), p(eo), p(im), None))
        self._done()
        return self.n(ra), self.n(rt), self.n(w), self.n(dp), self.n(im), self.n(we), self.n(de)

    def compact(self, rays_alive):
        N, p, torch = self.N, self.N.ptr, self.torch
        ra = self.i32(rays_alive)
        n = ra.shape[0]
        out = torch.full((max(n, 1),), -9, dtype=torch.int32, device=self.dev)
        cnt = torch.full((1,), -1, dtype=torch.int32, device=self.dev)
        N.check(self.lib.lnrf_compact_alive(p(ra), n, p(out), p(cnt), p(self._scratch_compact), self._scratch_compact.numel() * 8, None))
        self._done()
        assert int(self._scratch_compact.abs().sum().item()) == 0, "compact scratch not left zeroed"
        k = int(cnt.item())
        return self.n(out)[:k], k

    def grid_level_scales(self, L, per_level_scale, H):
        N, p = self.N, self.N.ptr
        out = self.torch.empty(L, device=self.dev)
        N.check(self.lib.lnrf_grid_level_scales(L, float(np.log2(per_level_scale)), H, p(out), None))
        self._done()
        return self.n(out)

    def grid_fwd(self, inputs, emb, offsets, per_level_scale, H, half=False, dy_dx=False, gridtype=0, align=False, interp=0,
                 scales=None, layout=1):
        N, p, torch = self.N, self.N.ptr, self.torch
        x = self.f32(inputs)
        e = self.f16(emb) if half else self.f32(emb)
        off = torch.from_numpy(np.asarray(offsets, np.int32))
        B, D = x.shape
        L, C = off.shape[0] - 1, e.shape[1]
        out = torch.full((B, L * C) if layout == 1 else (L, B, C), float("nan"), dtype=e.dtype, device=self.dev)
        dd = torch.full((B, L * D * C), float("nan"), dtype=e.dtype, device=self.dev) if dy_dx else None
        N.check(self.lib.lnrf_grid_encode_forward(p(x), p(e), p(off), p(out), B, D, C, L, float(np.log2(per_level_scale)), H, p(dd),
                                                  gridtype, int(align), interp, N.F16 if half else N.F32, layout, None))
        self._done()
        return (self.n(out), self.n(dd)) if dy_dx else self.n(out)

    def grid_bwd(self, grad, inputs, offsets, C, per_level_scale, H, half=False, gridtype=0, align=False, interp=0, scales=None,
                 layout=1, dy_dx=None):
        N, p, torch = self.N, self.N.ptr, self.torch
        x = self.f32(inputs)
        g = self.f16(grad) if half else self.f32(grad)
        off = torch.from_numpy(np.asarray(offsets, np.int32))
        B, D = x.shape
        L = off.shape[0] - 1
        ge = torch.zeros(int(offsets[-1]), C, dtype=g.dtype, device=self.dev)
        dd = None if dy_dx is None else (self.f16(dy_dx) if half else self.f32(dy_dx))
        gi = torch.zeros(B, D, dtype=g.dtype, device=self.dev) if dd is not None else None
        N.check(self.lib.lnrf_grid_encode_backward(p(g), p(x), None, p(off), p(ge), B, D, C, L, float(np.log2(per_level_scale)), H,
                                                   p(dd), p(gi), gridtype, int(align), interp, N.F16 if half else N.F32, layout, None))
        self._done()
        return (self.n(ge), self.n(gi)) if gi is not None else self.n(ge)

    def ffmlp_fwd(self, inputs, weights, in_dim, out_dim, hidden, n_layers, act=0, out_act=6, inference=False):
        N, p, torch = self.N, self.N.ptr, self.torch
        x, w = self.f16(inputs), self.f16(weights)
        B = x.shape[0]
        out = torch.full((B, out_dim), float("nan"), dtype=torch.float16, device=self.dev)
        if inference:
            N.check(self.lib.lnrf_ffmlp_inference(p(x), p(w), B, in_dim, out_dim, hidden, n_layers, act, out_act, None, p(out), None))
            self._done()
            return self.n(out)
        fb = torch.full((n_layers, B, hidden), float("nan"), dtype=torch.float16, device=self.dev)
        N.check(self.lib.lnrf_ffmlp_forward(p(x), p(w), B, in_dim, out_dim, hidden, n_layers, act, out_act, p(fb), p(out), None))
        self._done()
        return self.n(out), self.n(fb)

    def ffmlp_bwd(self, grad, inputs, weights, fwd_buf, in_dim, out_dim, hidden, n_layers, act=0, calc_grad_inputs=False):
        N, p, torch = self.N, self.N.ptr, self.torch
        g, x, w, fb = self.f16(grad), self.f16(inputs), self.f16(weights), self.f16(fwd_buf)
        B = x.shape[0]
        gw = torch.full((w.shape[0],), float("nan"), dtype=torch.float16, device=self.dev)
        gi = torch.full((B, in_dim), float("nan"), dtype=torch.float16, device=self.dev) if calc_grad_inputs else None
        nbytes = self.lib.lnrf_ffmlp_wgrad_scratch_bytes(in_dim, out_dim, hidden, n_layers)
        sc = torch.full((nbytes // 4,), float("nan"), device=self.dev)
        N.check(self.lib.lnrf_ffmlp_backward(p(g), p(x), p(w), p(fb), B, in_dim, out_dim, hidden, n_layers, act, 6,
                                             int(calc_grad_inputs), None, p(gi), p(gw), p(sc), nbytes, None))
        self._done()
        return self.n(gw), self.n(gi)

    def sh(self, dirs, degree, half=False):
        N, p, torch = self.N, self.N.ptr, self.torch
        d = self.f32(dirs).view(-1, 3)
        out = torch.full((d.shape[0], degree * degree), float("nan"), dtype=torch.float16 if half else torch.float32, device=self.dev)
        N.check(self.lib.lnrf_sh_encode_forward(p(d), p(out), d.shape[0], degree, None, N.F16 if half else N.F32, None))
        self._done()
        return self.n(out)


# =============================================================================================================
class RefBackend(_TorchDevice):
    """The reference's own CUDA extensions (oracle/_ref/*, built untouched from /root/reference by
    oracle/build_ref.py).  Calls mirror the reference wrappers' allocation contracts (torch.zeros where they do)."""
    name = "reference"

    def __init__(self):
        super().__init__()
        import importlib
        ref = os.path.join(ROOT, "oracle", "_ref")
        self.m = {}
        for mod in ("_raymarching", "_gridencoder", "_ffmlp", "_shencoder"):
            sys.path.insert(0, os.path.join(ref, mod))
            self.m[mod] = importlib.import_module(mod)
        self.rm, self.ge, self.ff, self.shm = (self.m[k] for k in ("_raymarching", "_gridencoder", "_ffmlp", "_shencoder"))
        self.ff.allocate_splitk(8)

    def _done(self):
        self.torch.cuda.synchronize()

    def near_far(self, rays_o, rays_d, aabb, min_near):
        o, d, a = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3), self.f32(aabb)
        n = o.shape[0]
        nears, fars = self.torch.empty(n, device=self.dev), self.torch.empty(n, device=self.dev)
        self.rm.near_far_from_aabb(o, d, a, n, float(min_near), nears, fars)
        self._done()
        return self.n(nears), self.n(fars)

    def morton3D(self, coords):
        c = self.i32(coords).view(-1, 3)
        out = self.torch.empty(c.shape[0], dtype=self.torch.int32, device=self.dev)
        self.rm.morton3D(c, c.shape[0], out)
        self._done()
        return self.n(out)

    def morton3D_invert(self, idx):
        c = self.i32(idx).view(-1)
        out = self.torch.empty(c.shape[0], 3, dtype=self.torch.int32, device=self.dev)
        self.rm.morton3D_invert(c, c.shape[0], out)
        self._done()
        return self.n(out)

    def packbits(self, grid, thresh):
        g = self.f32(grid).view(-1)
        n = g.shape[0] // 8
        out = self.torch.empty(n, dtype=self.torch.uint8, device=self.dev)
        self.rm.packbits(g, n, float(thresh), out)
        self._done()
        return self.n(out)

    def march_train(self, rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter=None):
        torch = self.torch
        o, d = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3)
        n = o.shape[0]
        g, ne, fa, no = self.u8(bitfield), self.f32(nears), self.f32(fars), self.f32(noises)
        xyzs, dirs, deltas = torch.zeros(M, 3, device=self.dev), torch.zeros(M, 3, device=self.dev), torch.zeros(M, 2, device=self.dev)
        rays = torch.empty(n, 3, dtype=torch.int32, device=self.dev)
        cnt = self.i32(np.zeros(2) if counter is None else counter)
        self.rm.march_rays_train(o, d, g, float(bound), float(dt_gamma), int(max_steps), n, int(C), int(H), M, ne, fa, xyzs, dirs,
                                 deltas, rays, cnt, no)
        self._done()
        return self.n(xyzs), self.n(dirs), self.n(deltas), self.n(rays), self.n(cnt)

    def composite_train_fwd(self, sigmas, rgbs, deltas, rays, T):
        torch = self.torch
        s, c, dl, r = self.f32(sigmas), self.f32(rgbs), self.f32(deltas), self.i32(rays)
        M, n = s.shape[0], r.shape[0]
        ws, depth, image = torch.empty(n, device=self.dev), torch.empty(n, device=self.dev), torch.empty(n, 3, device=self.dev)
        self.rm.composite_rays_train_forward(s, c, dl, r, M, n, float(T), ws, depth, image)
        self._done()
        return self.n(ws), self.n(depth), self.n(image)

    def composite_train_bwd(self, gws, gimg, sigmas, rgbs, deltas, rays, ws, image, T):
        torch = self.torch
        s, c, dl, r = self.f32(sigmas), self.f32(rgbs), self.f32(deltas), self.i32(rays)
        M, n = s.shape[0], r.shape[0]
        gs, gc = torch.zeros(M, device=self.dev), torch.zeros(M, 3, device=self.dev)
        self.rm.composite_rays_train_backward(self.f32(gws), self.f32(gimg), s, c, dl, r, self.f32(ws), self.f32(image), M, n, float(T),
                                              gs, gc)
        self._done()
        return self.n(gs), self.n(gc)

    def march(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, M_rows,
              dt_gamma=0.0, max_steps=1024, edit_bitfield=None):
        torch = self.torch
        o, d = self.f32(rays_o).view(-1, 3), self.f32(rays_d).view(-1, 3)
        ra, rt, g, ne, fa, no = self.i32(rays_alive), self.f32(rays_t), self.u8(bitfield), self.f32(nears), self.f32(fars), self.f32(noises)
        xyzs, dirs, deltas = (torch.zeros(M_rows, 3, device=self.dev), torch.zeros(M_rows, 3, device=self.dev),
                              torch.zeros(M_rows, 2, device=self.dev))
        if edit_bitfield is None:
            self.rm.march_rays(n_alive, n_step, ra, rt, o, d, float(bound), float(dt_gamma), int(max_steps), int(C), int(H), g, ne, fa,
                               xyzs, dirs, deltas, no)
            self._done()
            return self.n(xyzs), self.n(dirs), self.n(deltas)
        occ = torch.zeros(M_rows, dtype=torch.bool, device=self.dev)
        self.rm.march_rays_distill(n_alive, n_step, ra, rt, o, d, float(bound), float(dt_gamma), int(max_steps), int(C), int(H), g,
                                   self.u8(edit_bitfield), ne, fa, xyzs, dirs, deltas, occ, no)
        self._done()
        return self.n(xyzs), self.n(dirs), self.n(deltas), self.n(occ).astype(np.uint8)

    def composite(self, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, ws, depth, image, T, wes=None, depth_edit=None,
                  edit_occ=None):
        ra, rt = self.i32(rays_alive), self.f32(rays_t)
        s, c, dl = self.f32(sigmas), self.f32(rgbs), self.f32(deltas)
        w, dp, im = self.f32(ws), self.f32(depth), self.f32(image)
        if edit_occ is None:
            self.rm.composite_rays(n_alive, n_step, float(T), ra, rt, s, c, dl, w, dp, im)
            self._done()
            return self.n(ra), self.n(rt), self.n(w), self.n(dp), self.n(im)
        we, de = self.f32(wes), self.f32(depth_edit)
        eo = self.u8(edit_occ).bool()
        self.rm.composite_rays_distill(n_alive, n_step, float(T), ra, rt, s, c, dl, w, we, dp, de, eo, im)
        self._done()
        return self.n(ra), self.n(rt), self.n(w), self.n(dp), self.n(im), self.n(we), self.n(de)

    def grid_fwd(self, inputs, emb, offsets, per_level_scale, H, half=False, dy_dx=False, gridtype=0, align=False, interp=0,
                 scales=None):
        torch = self.torch
        x = self.f32(inputs)
        e = self.f16(emb) if half else self.f32(emb)
        off = self.i32(offsets)
        B, D = x.shape
        L, C = off.shape[0] - 1, e.shape[1]
        out = torch.empty(L, B, C, dtype=e.dtype, device=self.dev)
        dd = torch.empty(B, L * D * C, dtype=e.dtype, device=self.dev) if dy_dx else None
        self.ge.grid_encode_forward(x, e, off, out, B, D, C, L, float(np.log2(per_level_scale)), H, dd, gridtype, bool(align), interp)
        self._done()
        out = out.permute(1, 0, 2).reshape(B, L * C)  # grid.py:57
        return (self.n(out), self.n(dd)) if dy_dx else self.n(out)

    def grid_bwd(self, grad, inputs, offsets, C, per_level_scale, H, half=False, gridtype=0, align=False, interp=0, scales=None):
        torch = self.torch
        x = self.f32(inputs)
        g = self.f16(grad) if half else self.f32(grad)
        off = self.i32(offsets)
        B, D = x.shape
        L = off.shape[0] - 1
        g = g.view(B, L, C).permute(1, 0, 2).contiguous()  # grid.py:75
        e = torch.zeros(int(offsets[-1]), C, dtype=g.dtype, device=self.dev)
        ge = torch.zeros_like(e)
        self.ge.grid_encode_backward(g, x, e, off, ge, B, D, C, L, float(np.log2(per_level_scale)), H, None, None, gridtype, bool(align),
                                     interp)
        self._done()
        return self.n(ge)

    def ffmlp_fwd(self, inputs, weights, in_dim, out_dim, hidden, n_layers, act=0, out_act=6):
        torch = self.torch
        x, w = self.f16(inputs), self.f16(weights)
        B = x.shape[0]
        out = torch.empty(B, out_dim, dtype=torch.float16, device=self.dev)
        fb = torch.empty(n_layers, B, hidden, dtype=torch.float16, device=self.dev)
        self.ff.ffmlp_forward(x, w, B, in_dim, out_dim, hidden, n_layers, act, out_act, fb, out)
        self._done()
        return self.n(out), self.n(fb)

    def ffmlp_bwd(self, grad, inputs, weights, fwd_buf, in_dim, out_dim, hidden, n_layers, act=0, calc_grad_inputs=False):
        torch = self.torch
        g, x, w, fb = self.f16(grad), self.f16(inputs), self.f16(weights), self.f16(fwd_buf)
        B = x.shape[0]
        gi = torch.zeros_like(x) if calc_grad_inputs else torch.zeros(1, dtype=torch.float16, device=self.dev)
        gw = torch.zeros_like(w)
        bb = torch.zeros(n_layers, B, hidden, dtype=torch.float16, device=self.dev)
        self.ff.ffmlp_backward(g, x, w, fb, B, in_dim, out_dim, hidden, n_layers, act, 6, bool(calc_grad_inputs), bb, gi, gw)
        self._done()
        return self.n(gw), (self.n(gi) if calc_grad_inputs else None)

    def sh(self, dirs, degree):
        d = self.f32(dirs).view(-1, 3)
        out = self.torch.empty(d.shape[0], degree * degree, device=self.dev)
        self.shm.sh_encode_forward(d, out, d.shape[0], 3, degree, None)
        self._done()
        return self.n(out)
