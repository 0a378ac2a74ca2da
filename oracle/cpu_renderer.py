"""CPU baseline: a PyTorch-CPU restatement of the reference's NON-cuda_ray renderer.  TEST/BENCH INFRASTRUCTURE ONLY
(used by bench.py's `cpu_baseline` leg and `--impl reference`; never imported by the laenerf_b200 package).

BASELINE.json configs[0] / SURVEY.md section 3.3: `NeRFRenderer.run` (nerf/renderer.py:128-256) driving the nn.Linear
`NeRFNetwork` of nerf/network.py:10-124 with frequency encodings, fp32, on the host cores.  The path is not
CPU-runnable as shipped (it calls the CUDA-only raymarching.near_far_from_aabb / freqencoder / shencoder), so the
three substitutions SURVEY.md 3.3 lists are made and nothing else: the slab test in torch (raymarching.cu:108-144),
the pure-torch FreqEncoder of encoding.py:5-43 for positions and for directions.  /root/reference does not exist
on the GPU box, hence a restatement ("kind": "port") instead of an import.
"""
from __future__ import annotations

import math
import os
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class FreqEncoder(nn.Module):  # encoding.py:5-43 with get_encoder's multires=6: include_input, log-sampled bands
    def __init__(self, input_dim=3, multires=6):
        super().__init__()
        self.freq_bands = (2.0 ** torch.linspace(0.0, multires - 1, multires)).tolist()
        self.output_dim = input_dim + input_dim * multires * 2

    def forward(self, x, **kwargs):
        out = [x]
        for f in self.freq_bands:
            out.append(torch.sin(x * f))
            out.append(torch.cos(x * f))
        return torch.cat(out, dim=-1)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near):  # raymarching.cu:108-144
    rd = 1.0 / rays_d
    t0 = (aabb[:3] - rays_o) * rd
    t1 = (aabb[3:] - rays_o) * rd
    near = torch.minimum(t0, t1).amax(-1)
    far = torch.maximum(t0, t1).amin(-1)
    miss = near > far
    near = near.clamp(min=min_near)
    big = torch.finfo(torch.float32).max
    return torch.where(miss, torch.full_like(near, big), near), torch.where(miss, torch.full_like(far, big), far)


class CpuNeRF(nn.Module):
    """network.py:10-124 (frequency / frequency encodings, no background model) + renderer.py:128-256."""

    def __init__(self, bound=1.0, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                 min_near=0.2, density_scale=1.0):
        super().__init__()
        self.bound, self.min_near, self.density_scale, self.geo_feat_dim = bound, min_near, density_scale, geo_feat_dim
        self.register_buffer("aabb", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))
        self.encoder, self.encoder_dir = FreqEncoder(3, 6), FreqEncoder(3, 6)
        dims = [self.encoder.output_dim] + [hidden_dim] * (num_layers - 1) + [1 + geo_feat_dim]
        self.sigma_net = nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:])])
        dims = [self.encoder_dir.output_dim + geo_feat_dim] + [hidden_dim_color] * (num_layers_color - 1) + [3]
        self.color_net = nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:])])

    def density(self, x):  # network.py:126-145
        h = self.encoder(x, bound=self.bound)
        for l, lin in enumerate(self.sigma_net):
            h = lin(h)
            if l != len(self.sigma_net) - 1:
                h = F.relu(h, inplace=True)
        return torch.exp(h[..., 0]), h[..., 1:]  # trunc_exp forward (activation.py:9-12)

    def color(self, d, mask, geo_feat):  # network.py:160-190: only where mask, zeros elsewhere
        rgbs = torch.zeros(mask.shape[0], 3, dtype=geo_feat.dtype)
        if mask.any():
            h = torch.cat([self.encoder_dir(d[mask]), geo_feat[mask]], dim=-1)
            for l, lin in enumerate(self.color_net):
                h = lin(h)
                if l != len(self.color_net) - 1:
                    h = F.relu(h, inplace=True)
            rgbs[mask] = torch.sigmoid(h)
        return rgbs

    def run(self, rays_o, rays_d, num_steps=512, bg_color=1, perturb=False):  # renderer.py:128-256, upsample_steps=0
        N = rays_o.shape[0]
        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb, self.min_near)
        nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)
        z_vals = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand(N, num_steps)
        z_vals = nears + (fars - nears) * z_vals
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z_vals = z_vals + (torch.rand(z_vals.shape) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, self.aabb[:3]), self.aabb[3:])
        sigma, geo = self.density(xyzs.reshape(-1, 3))
        sigma = sigma.view(N, num_steps)
        deltas = z_vals[..., 1:] - z_vals[..., :-1]
        deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * self.density_scale * sigma)
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        mask = weights > 1e-4
        rgbs = self.color(dirs.reshape(-1, 3), mask.reshape(-1), geo).view(N, -1, 3)
        weights_sum = weights.sum(dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
        return image + (1 - weights_sum).unsqueeze(-1) * bg_color


def time_train_steps(rays_o, rays_d, gt, steps=3, warmup=1, num_steps=512, bound=1.0, min_near=0.2, threads=None):
    """fwd + bwd + Adam (main_nerf.py:223) per step on the host cores; returns (seconds per step, threads used)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = CpuNeRF(bound=bound, min_near=min_near)
    # the slab test returns FLT_MAX for rays that miss the box; the reference's `run` would produce NaNs for them
    # (inf - inf), so -- as in real training images -- only rays that hit the bound-1 box are used
    near, far = near_far_from_aabb(rays_o, rays_d, model.aabb, min_near)
    hit = near < far
    rays_o, rays_d, gt = rays_o[hit], rays_d[hit], gt[hit]
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        image = model.run(rays_o, rays_d, num_steps=num_steps, bg_color=1, perturb=True)
        loss = F.mse_loss(image, gt)
        loss.backward()
        opt.step()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads, int(hit.sum())
