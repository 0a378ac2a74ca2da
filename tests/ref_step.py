"""TEST INFRASTRUCTURE: one NeRF training step driven through the REFERENCE's own CUDA extensions (oracle/_ref/*, built
untouched from /root/reference by oracle/build_ref.py) the way the reference's Python wrappers drive them, so that
the GPU tier can (a) check a whole step of laenerf_b200 against a whole step of the reference and (b) time the
reference step on the same B200 (tests/bench_gpu_reference.py -> profiles/*_gpu_reference.json).

What each wrapper allocates / copies is restated from the reference wrappers (cited per class); nothing here is
imported by the laenerf_b200 package or by bench.py's product path.
"""
from __future__ import annotations

import importlib
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_mods = {}


def ref_available() -> bool:
    return all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", m, m + ".so"))
               for m in ("_raymarching", "_gridencoder", "_ffmlp", "_shencoder"))


def backend(name: str):
    if name not in _mods:
        d = os.path.join(ROOT, "oracle", "_ref", name)
        if d not in sys.path:
            sys.path.insert(0, d)
        _mods[name] = importlib.import_module(name)
        if name == "_ffmlp":
            _mods[name].allocate_splitk(8)
    return _mods[name]


class _march_train(Function):  # raymarching/raymarching.py:161-235
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, bitfield, C, H, nears, fars, counter, mean_count, perturb, align, force_all_rays, dt_gamma, max_steps):
        rm = backend("_raymarching")
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        dev = rays_o.device
        xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        noises = torch.rand(N, device=dev) if perturb else torch.zeros(N, device=dev)
        rm.march_rays_train(rays_o, rays_d, bitfield, float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), M, nears, fars,
                            xyzs, dirs, deltas, rays, counter, noises)
        if force_all_rays or mean_count <= 0:
            m = int(counter[0].item())
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
            torch.cuda.empty_cache()
        return xyzs, dirs, deltas, rays


class _composite_train(Function):  # raymarching/raymarching.py:238-291
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh):
        rm = backend("_raymarching")
        sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, float(T_thresh), ws, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, depth, image)
        ctx.dims = [M, N, T_thresh]
        return ws, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g_ws, g_depth, g_image):
        rm = backend("_raymarching")
        g_ws, g_image = g_ws.contiguous(), g_image.contiguous()
        sigmas, rgbs, deltas, rays, ws, depth, image = ctx.saved_tensors
        M, N, T = ctx.dims
        gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        rm.composite_rays_train_backward(g_ws, g_image, sigmas, rgbs, deltas, rays, ws, image, M, N, float(T), gs, gc)
        return gs, gc, None, None, None


class _grid(Function):  # gridencoder/grid.py:24-89
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, H):
        ge = backend("_gridencoder")
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        ge.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, None, 0, False, 0)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dims = [B, D, C, L, S, H]
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        ge = backend("_gridencoder")
        inputs, embeddings, offsets = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        g_emb = torch.zeros_like(embeddings)
        ge.grid_encode_backward(grad, inputs, embeddings, offsets, g_emb, B, D, C, L, S, H, None, None, 0, False, 0)
        return None, g_emb, None, None, None


class _ffmlp(Function):  # ffmlp/ffmlp.py:15-83
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.half)
    def forward(ctx, inputs, weights, in_dim, out_dim, hidden, n_layers, calc_grad_inputs):
        ff = backend("_ffmlp")
        B = inputs.shape[0]
        inputs, weights = inputs.contiguous(), weights.contiguous()
        outputs = torch.empty(B, out_dim, device=inputs.device, dtype=inputs.dtype)
        fb = torch.empty(n_layers, B, hidden, device=inputs.device, dtype=inputs.dtype)
        ff.ffmlp_forward(inputs, weights, B, in_dim, out_dim, hidden, n_layers, 0, 6, fb, outputs)
        ctx.save_for_backward(inputs, weights, outputs, fb)
        ctx.dims = (in_dim, out_dim, hidden, n_layers, calc_grad_inputs)
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        ff = backend("_ffmlp")
        B = grad.shape[0]
        grad = grad.contiguous()
        inputs, weights, outputs, fb = ctx.saved_tensors
        in_dim, out_dim, hidden, n_layers, cgi = ctx.dims
        gi = torch.zeros_like(inputs) if cgi else torch.zeros(1, device=grad.device, dtype=grad.dtype)
        gw = torch.zeros_like(weights)
        bb = torch.zeros(n_layers, B, hidden, device=grad.device, dtype=grad.dtype)
        ff.ffmlp_backward(grad, inputs, weights, fb, B, in_dim, out_dim, hidden, n_layers, 0, 6, cgi, bb, gi, gw)
        return (gi if cgi else None), gw, None, None, None, None, None


def _ffmlp_module_forward(x, weights, in_dim, out_dim, n_layers):  # ffmlp/ffmlp.py:147-168 (pad to 128, slice)
    B, C = x.shape
    pad = 128 - (B % 128)
    if pad > 0:
        x = torch.cat([x, torch.zeros(pad, C, dtype=x.dtype, device=x.device)], dim=0)
    out = _ffmlp.apply(x, weights, in_dim, 16, 64, n_layers, x.requires_grad)
    return out[:B, :out_dim]


class _sh(Function):  # shencoder/sphere_harmonics.py:14-58
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree):
        sh = backend("_shencoder")
        inputs = inputs.contiguous()
        B = inputs.shape[0]
        out = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
        sh.sh_encode_forward(inputs, out, B, 3, degree, None)
        return out


class _trunc_exp(Function):  # activation.py:5-17
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class RefNeRF(nn.Module):
    """network_ff.NeRFNetwork + the cuda_ray training branch of NeRFRenderer.run_cuda (renderer.py:284-333) on the
    reference's extensions.  Parameters/buffers are copied from a laenerf_b200.nerf.NeRFNetwork so both see the same state."""

    def __init__(self, ours):
        super().__init__()
        self.bound, self.cascade, self.grid_size = ours.bound, ours.cascade, ours.grid_size
        self.min_near, self.density_scale = ours.min_near, ours.density_scale
        self.per_level_scale, self.base_resolution = ours.encoder.per_level_scale, ours.encoder.base_resolution
        self.embeddings = nn.Parameter(ours.encoder.embeddings.detach().clone())
        self.register_buffer("offsets", ours.encoder.offsets.clone())
        self.w_sigma = nn.Parameter(ours.sigma_net.weights.detach().clone())
        self.w_color = nn.Parameter(ours.color_net.weights.detach().clone())
        self.ns, self.nc = ours.sigma_net.num_layers, ours.color_net.num_layers
        self.register_buffer("aabb", ours.aabb_train.clone())
        self.register_buffer("density_bitfield", ours.density_bitfield.clone())
        self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
        self.mean_count, self.local_step = 0, 0

    def network(self, x, d):  # network_ff.py:51-79
        x = (x + self.bound) / (2 * self.bound)
        x = _grid.apply(x, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution)
        h = _ffmlp_module_forward(x, self.w_sigma, 32, 16, self.ns)
        sigma = _trunc_exp.apply(h[..., 0])
        geo = h[..., 1:]
        d = _sh.apply(d, 4)
        p = torch.zeros_like(geo[..., :1])
        h = torch.cat([d, geo, p], dim=-1)
        h = _ffmlp_module_forward(h, self.w_color, 32, 3, self.nc)
        return sigma, torch.sigmoid(h)

    def render_train(self, rays_o, rays_d, bg_color=1, perturb=True):
        rm = backend("_raymarching")
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        n = rays_o.shape[0]
        nears, fars = torch.empty(n, device=rays_o.device), torch.empty(n, device=rays_o.device)
        rm.near_far_from_aabb(rays_o, rays_d, self.aabb, n, float(self.min_near), nears, fars)
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        xyzs, dirs, deltas, rays = _march_train.apply(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size,
                                                      nears, fars, counter, self.mean_count, perturb, 128, False, 0, 1024)
        sigmas, rgbs = self.network(xyzs, dirs)
        sigmas = self.density_scale * sigmas
        ws, depth, image = _composite_train.apply(sigmas, rgbs, deltas, rays, 1e-4)
        image = image + (1 - ws).unsqueeze(-1) * bg_color
        return image, xyzs.shape[0]

    def update_mean_count(self):
        total = min(16, self.local_step)
        if total > 0:
            self.mean_count = int(self.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0


class RefTrainStep:
    """Trainer.train_step + the `-O` optimizer recipe (nerf/utils.py:535-642, 1474-1484; main_nerf.py:223)."""

    def __init__(self, model: RefNeRF, lr=1e-2):
        self.model = model
        self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
        self.scaler = torch.amp.GradScaler("cuda")

    def __call__(self, rays_o, rays_d, gt):
        self.optimizer.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            image, m = self.model.render_train(rays_o, rays_d)
            loss = torch.nn.functional.mse_loss(image, gt, reduction="none").mean(-1).mean()
        self.scaler.scale(loss).backward()
        self.scaler.step(self.optimizer)
        self.scaler.update()
        return loss, m
