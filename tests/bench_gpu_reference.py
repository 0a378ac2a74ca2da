#!/usr/bin/env python
"""TEST-TIER measurement (not bench.py's product path): the reference's OWN CUDA extensions (oracle/_ref, built for
sm_100a from the untouched sources) driven through tests/ref_step.py on the same B200, same synthetic lego-shape
scene, same 4096-ray batches as bench.py -- the comparator for BASELINE.json's ">= 10x the reference's own extensions
on one B200 for the lego-shape 800x800 training step" target.  Prints one JSON line; run under gpurun.

    python tests/bench_gpu_reference.py [--steps 30] [--warmup 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import ref_step
    from laenerf_b200.nerf import GraphedTrainStep, NeRFNetwork, TrainStep
    from laenerf_b200.scene import get_rays_np, make_scene

    if not ref_step.ref_available():
        print(json.dumps({"gpu_reference": None, "why": "oracle/_ref not built (python oracle/build_ref.py)"}))
        return
    dev = torch.device("cuda", 0)
    sc = make_scene("lego", seed=0, n_poses=8)
    torch.manual_seed(0)
    ours = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_thresh=sc.density_thresh).to(dev)
    ours.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    ref = ref_step.RefNeRF(ours).to(dev)
    rng = np.random.default_rng(1000)
    batches = []
    for b in range(8):
        ro, rd, _ = get_rays_np(sc.poses[b % len(sc.poses)], sc.intrinsics, sc.H, sc.W, N=4096, rng=rng)
        gt = rng.random((4096, 3), dtype=np.float32)
        batches.append(tuple(torch.from_numpy(x).to(dev) for x in (ro, rd, gt)))

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- reference extensions: eager, as the reference's trainer issues them -----------------------------------------
    rstep = ref_step.RefTrainStep(ref)
    pts = []
    for i in range(args.warmup):
        rstep(*batches[i % 8])
        if i == 0:
            ref.update_mean_count()
    ref.update_mean_count()
    ref_ms = timed(lambda i: pts.append(rstep(*batches[i % 8])[1]), args.steps)

    # per-kernel view of the same step: CUDA events around each reference backend call (extension time only)
    kern = {}
    import functools
    for modname in ("_raymarching", "_gridencoder", "_ffmlp", "_shencoder"):
        mod = ref_step.backend(modname)
        for fn in ("near_far_from_aabb", "march_rays_train", "composite_rays_train_forward", "composite_rays_train_backward",
                   "grid_encode_forward", "grid_encode_backward", "ffmlp_forward", "ffmlp_backward", "sh_encode_forward"):
            if hasattr(mod, fn):
                orig = getattr(mod, fn)

                def wrapped(*a, _o=orig, _n=fn):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    r = _o(*a)
                    e1.record()
                    kern.setdefault(_n, []).append((e0, e1))
                    return r
                setattr(mod, fn, wrapped)
    for i in range(5):
        rstep(*batches[i % 8])
    torch.cuda.synchronize()
    kernels = {k: {"calls_per_step": len(v) / 5, "mean_ms": sum(a.elapsed_time(b) for a, b in v) / len(v)} for k, v in kern.items()}
    ext_ms = sum(sum(a.elapsed_time(b) for a, b in v) for v in kern.values()) / 5

    # ---- ours on the same batches: eager modules and the graphed product path ---------------------------------------
    step = TrainStep(ours)
    for i in range(args.warmup):
        step(*batches[i % 8])
        if i == 0:
            ours.update_mean_count()
    ours.update_mean_count()
    ours_eager_ms = timed(lambda i: step(*batches[i % 8]), args.steps)
    ours.update_mean_count()
    g = GraphedTrainStep(step, 4096)
    g.capture(*batches[0])
    for i in range(3):
        g(*batches[i % 8])
    ours_graph_ms = timed(lambda i: g(*batches[i % 8]), args.steps)

    line = {
        "what": "lego-shape 800x800 hash-grid NeRF training step, 4096 rays, fp16 autocast, Adam -- reference extensions vs laenerf_b200 on the same GPU",
        "gpu_reference": {"ms_per_step": ref_ms, "rays_per_s": 4096 / (ref_ms * 1e-3), "samples_per_step_padded": int(np.mean(pts)),
                          "mode": "eager (the reference's wrappers + torch Adam/GradScaler, restated in tests/ref_step.py)",
                          "extension_kernels_ms_per_step": ext_ms, "kernels": kernels},
        "ours": {"eager_ms_per_step": ours_eager_ms, "graph_ms_per_step": ours_graph_ms, "rays_per_s": 4096 / (ours_graph_ms * 1e-3),
                 "fused_optimizer": step.fused_optimizer},
        "speedup_vs_reference_step": ref_ms / ours_graph_ms,
        "speedup_vs_reference_step_eager": ref_ms / ours_eager_ms,
        "gpu": torch.cuda.get_device_name(0),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
