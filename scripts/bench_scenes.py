#!/usr/bin/env python
"""The other BASELINE.json configs on one GPU, as evidence beside bench.py's headline line (which stays on configs[1]):

  configs[2]  flower-shape 504x378 (bound 2, 2 cascades): training step, graph replay
  configs[3]  bonsai-shape 779x519 (bound 16, 5 cascades): one full test view through the device-driven render loop
  configs[4]  edit stage on the flower shape: distillation render of one view against an edit grid (run_cuda_distill),
              then StyleTrainStep iterations (LAENeRF style network: its own hash grid + two MLPs + palette) on the masked points

Prints one JSON object; every number is CUDA-event time with warm-up, inputs resident on the device.
    python scripts/bench_scenes.py > profiles/<tag>_scenes.json
"""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from laenerf_b200.nerf import GraphedTrainStep, NeRFNetwork, TrainStep
from laenerf_b200.scene import get_rays_np, make_scene
from laenerf_b200.style_encoder import LAENeRF, StyleTrainStep

dev = torch.device("cuda", 0)
N_RAYS = 4096


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build(name):
    sc = make_scene(name, seed=0, n_poses=8)
    torch.manual_seed(0)
    model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_thresh=sc.density_thresh).to(dev)
    model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    return sc, model


def train_case(name):
    sc, model = build(name)
    step = TrainStep(model)
    rng = np.random.default_rng(1)
    batches = []
    for b in range(4):
        ro, rd, _ = get_rays_np(sc.poses[b % len(sc.poses)], sc.intrinsics, sc.H, sc.W, N=N_RAYS, rng=rng)
        batches.append(tuple(torch.from_numpy(x).to(dev) for x in (ro, rd, rng.random((N_RAYS, 3), dtype=np.float32))))
    for i in range(5):
        step(*batches[i % 4])
        if i == 0:
            model.update_mean_count()
    model.update_mean_count()
    g = GraphedTrainStep(step, N_RAYS)
    g.capture(*batches[0])
    k = [0]

    def one():
        g(*batches[k[0] % 4])
        k[0] += 1
    ms = timed(one, 30)
    samples = int(model.step_counter[:16, 0].float().mean().item())
    return dict(scene=name, bound=sc.bound, cascades=model.cascade, image=[sc.H, sc.W], rays=N_RAYS, ms_per_step=ms, rays_per_s=N_RAYS / ms * 1e3,
                samples_per_step=samples, samples_per_ray=samples / N_RAYS, occupancy=sc.occupancy_fraction())


def render_case(name, schedule):
    sc, model = build(name)
    ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    model.eval()
    model.render_schedule = schedule
    out = {}

    def one():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out["o"] = model.render(ro, rd, perturb=False, bg_color=1)
    ms = timed(one, 3, warm=2)
    o = out["o"]
    return dict(scene=name, bound=sc.bound, cascades=model.cascade, image=[sc.H, sc.W], rays=int(ro.shape[0]), schedule=schedule, ms_per_frame=ms,
                rounds=o.get("rounds"), sample_slots=o["num_points"], slots_msamples_per_s=o["num_points"] / ms / 1e3,
                rays_per_s=ro.shape[0] / ms * 1e3, finite=bool(torch.isfinite(o["image"]).all()))


def edit_case():
    sc, model = build("flower")
    ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    model.eval()
    model.render_schedule = "fast"
    # edit grid: the occupied cells of the lower-x half of the first cascade (same Morton / cascade layout as the density bitfield)
    edit = model.density_bitfield.clone()
    edit[edit.numel() // 4: edit.numel() // 2] = 0
    edit[3 * edit.numel() // 4:] = 0
    out = {}

    def distill():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out["o"] = model.run_cuda_distill(ro, rd, edit, perturb=False)
    ms_distill = timed(distill, 3, warm=2)
    o = out["o"]
    mask = o["weights_edit_sum"] > 0.05
    x_term, d = o["x_term"][mask].contiguous(), rd[mask].contiguous()
    masked = int(x_term.shape[0])
    # the reference's fp16 regularisers overflow beyond ~75k points (sum of (1 - max w) in half, style_encoder.py:185-189): one
    # iteration trains on a view's worth of at most 49 152 masked points here
    x_term, d = x_term[:49152].contiguous(), d[:49152].contiguous()
    K = int(x_term.shape[0])
    params = SimpleNamespace(bound=sc.bound, num_palette_bases=8, style_weight=0.0, weight_loss_uniform=1e-6, weight_loss_non_uniform=1e-6,
                             offset_loss=1e-6, palette_loss_valid=1e-3, palette_loss_distinct=1e-3)
    torch.manual_seed(1)
    style = LAENeRF(params, dir_encoding="sphere_harmonics").to(dev)
    st = StyleTrainStep(style, params)
    target = torch.rand(K, 3, device=dev)
    losses = []

    def it():
        losses.append(st(x_term, d, target)[0])
    ms_style = timed(it, 20, warm=5)
    l = [float(x) for x in losses]
    return dict(scene="flower", image=[sc.H, sc.W], rays=int(ro.shape[0]), distill_ms_per_view=ms_distill, masked_points_of_view=masked, points_per_style_step=K,
                style_step_ms=ms_style, style_points_per_s=K / ms_style * 1e3, loss_first=l[0], loss_last=l[-1],
                note="run_cuda_distill (march_rays_distill / composite_rays_distill rounds) + StyleTrainStep (hash grid fwd/bwd, SH-3, two FFMLP nets, "
                     "palette mix, MSE + regularisers, GradScaler, torch Adam)")


res = {"gpu": torch.cuda.get_device_name(0),
       "train": [train_case("lego"), train_case("flower"), train_case("bonsai")],
       "render": [render_case("lego", "fast"), render_case("flower", "fast"), render_case("bonsai", "fast"), render_case("bonsai", "reference")],
       "edit_stage": edit_case()}
print(json.dumps(res, indent=1))
