#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-marched NeRF step (BASELINE.json: train rays/s & render Msamples/s on a
lego-shape 800x800 synthetic scene, % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); training shards RAYS (4096 per GPU, weak scaling) with one
all-reduce of the hash-grid + MLP gradients per step.  A "step" is one full training step of configs[1]:
near/far -> march_rays_train -> hash-grid encode -> sigma MLP -> SH -> colour MLP -> composite -> MSE -> backward of
all of it -> Adam, under fp16 autocast exactly as `-O` runs it.  One JSON line is printed by rank 0.

`--impl reference` times the CPU arm: the PyTorch-CPU restatement of the reference's non-cuda_ray renderer
(oracle/cpu_renderer.py, BASELINE.json configs[0]) on the host cores, same metric and unit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_RAYS = 4096
_OUT = sys.stdout
WORKLOAD = "lego-shape 800x800 hash-grid NeRF training step (16 levels, 2^19 table, 4096 rays/GPU, cuda_ray, fp16 autocast, Adam)"

# algorithmic bytes / FLOPs per unit (restated in DESIGN.md section 7)
ALGO = {
    "lnrf_march_rays_train": dict(bound="hbm", per_ray=48, per_sample=32),
    "lnrf_grid_encode_forward": dict(bound="hbm", per_ray=0, per_sample_padded=588),
    "lnrf_grid_encode_backward": dict(bound="hbm", per_ray=0, per_sample_padded=588),
    "lnrf_grid_encode_forward_world": dict(bound="hbm", per_ray=0, per_sample_padded=588),   # same kernels, world-coordinate inputs
    "lnrf_grid_encode_backward_world": dict(bound="hbm", per_ray=0, per_sample_padded=588),
    "lnrf_composite_rays_train_forward": dict(bound="hbm", per_ray=32, per_sample=24),
    "lnrf_composite_rays_train_backward": dict(bound="hbm", per_ray=44, per_sample=40),
    # row f-5: composite + blend + depth normalisation + MSE in one launch (gt 12 + image_raw 12 + nears/fars 8 B/ray on top)
    "lnrf_composite_loss_train_forward": dict(bound="hbm", per_ray=64, per_sample=24),
    "lnrf_composite_loss_train_backward": dict(bound="hbm", per_ray=60, per_sample=40),
    "lnrf_ffmlp_forward": dict(bound="tensor", flops_per_sample_padded=36864 / 2),   # mean of sigma (14336) and colour (22528) nets
    "lnrf_ffmlp_backward": dict(bound="tensor", flops_per_sample_padded=73728 / 2),
    "lnrf_sh_encode_forward": dict(bound="hbm", per_ray=0, per_sample_padded=12 + 64),
    "lnrf_near_far_from_aabb": dict(bound="hbm", per_ray=32, per_sample=0),
    # fused rows f-1 / f-4 (one C-ABI call each; nerf_backward = colour-net + sigma-net launches and two finalize launches)
    "lnrf_nerf_forward": dict(bound="tensor", flops_per_sample_padded=36864),
    "lnrf_nerf_backward": dict(bound="tensor", flops_per_sample_padded=73728),
    "lnrf_adam_step": dict(bound="hbm", per_param=30),            # g16 R+W 4, p/m/v R+W 24, p16 W 2
    "lnrf_grad_nonfinite_check": dict(bound="hbm", per_param=2),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]), tensor=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class TimedLib:
    """Proxy around the ctypes library that brackets selected entry points with CUDA events on the launching stream
    (per-kernel durations measured live inside the timed region, as the roofline contract asks)."""

    def __init__(self, lib, names, torch):
        self._lib, self._names, self._torch = lib, set(names), torch
        self.events = {n: [] for n in names}
        self.enabled = False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in self._names:
            return fn

        def wrapped(*a):
            if not self.enabled:
                return fn(*a)
            torch = self._torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            self.events[name].append((e0, e1))
            return r
        return wrapped

    def summary(self):
        out = {}
        for n, ev in self.events.items():
            if ev:
                ms = [a.elapsed_time(b) for a, b in ev]
                out[n] = dict(calls=len(ms), mean_ms=sum(ms) / len(ms), total_ms=sum(ms))
        return out


def run_reference(args, rank, world):
    """CPU arm: oracle/cpu_renderer.py (PyTorch-CPU port of NeRFRenderer.run + nn.Linear NeRFNetwork, freq encodings)."""
    if rank != 0:
        return
    import torch
    from cases import scene_rays
    from oracle import cpu_renderer
    n = min(N_RAYS, int(os.environ.get("LNRF_CPU_RAYS", "2048")))
    sc, ro, rd, rng = scene_rays("lego", n, 0)
    gt = torch.rand(n, 3, generator=torch.Generator().manual_seed(0))
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    sec, threads, used = cpu_renderer.time_train_steps(torch.from_numpy(ro), torch.from_numpy(rd), gt, steps=steps, warmup=warm,
                                                       num_steps=512, bound=sc.bound, min_near=sc.min_near)
    value = used / sec
    sample = f"{used} rays x 512 uniform samples/ray per step (hits of a {n}-ray draw), {steps} timed + {warm} warm-up steps, fp32"
    line = {"impl": "reference", "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "arm": "reference non-cuda_ray renderer (nerf/renderer.py run + nerf/network.py, "
                                            "frequency encodings) restated for CPU in oracle/cpu_renderer.py; bounded sample"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-render", action="store_true", help="skip the full-image render measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    # ONE JSON line on stdout, whatever the libraries print (NCCL writes its version banner to fd 1): keep the real stdout aside
    # and point fd 1 at stderr for the rest of the process
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # whatever NCCL logs (its version banner included) stays off stdout: rank 0 prints ONE JSON line
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from laenerf_b200 import _native
    from laenerf_b200.nerf import GraphedTrainStep, NeRFNetwork, TrainStep
    from laenerf_b200.parallel import gather_tiles, init_distributed, tile_shard_indices
    from laenerf_b200.scene import get_rays_np, make_scene

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timed = TimedLib(_native.lib(), list(ALGO), torch)
    _native._lib = timed

    # ---- synthetic lego-shape scene, random-init weights (seed = rank for the ray draws) -------------------------
    sc = make_scene("lego", seed=0, n_poses=8)
    torch.manual_seed(0)
    model = NeRFNetwork(bound=sc.bound, min_near=sc.min_near, density_thresh=sc.density_thresh).to(dev)
    model.set_density_grid(torch.from_numpy(sc.density_grid).to(dev), thresh=10.0)
    step = TrainStep(model, world_size=world)
    rng = np.random.default_rng(1000 + rank)
    n_batches = 8
    host_batches = []
    for b in range(n_batches):
        ro, rd, _ = get_rays_np(sc.poses[b % len(sc.poses)], sc.intrinsics, sc.H, sc.W, N=N_RAYS, rng=rng)
        gt = rng.random((N_RAYS, 3), dtype=np.float32)
        host_batches.append(tuple(torch.from_numpy(x).pin_memory() for x in (ro, rd, gt)))
    dev_batches = [tuple(x.to(dev) for x in hb) for hb in host_batches]

    # ---- warm-up: the first step sizes the sample buffer from the counters like the reference (renderer.py:643-647) ----
    for i in range(args.warmup):
        step(*dev_batches[i % n_batches])
        if i == 0:
            model.update_mean_count()
    model.update_mean_count()
    barrier()

    def timed_loop(fn, n):
        """n calls of fn(i) bracketed by barrier + synchronize, CUDA events, max over ranks -> total ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- pass 1 (eager, Python-issued launches): per-kernel CUDA events live on the launching stream ------------------
    points = []
    timed.enabled = True
    eager_l0 = _native.launch_count()
    eager_ms = timed_loop(lambda i: points.append(step(*dev_batches[i % n_batches])[1]["num_points"]), args.steps)
    timed.enabled = False
    eager_launches = _native.launch_count() - eager_l0
    actual = int(model.step_counter[: min(16, args.steps), 0].float().mean().item())
    model.update_mean_count()

    # ---- pass 2 (product path): the whole step replayed from one CUDA graph; device-resident inputs ----------------------
    # look-ahead (march of batch k+1 beside the exchange + Adam of batch k) paid while the exchange was an NCCL all-reduce (N = 2:
    # 0.623 -> 0.590 ms); with the fused peer-memory kernel it measures within +-1 % of the plain graph at N = 2 / 4 / 8
    # (profiles/r1k_bench_n*.json, r1l_bench_n8.json: 0.484 / 0.468 / 0.474 vs 0.480 / 0.462 / 0.479 ms), and on one GPU the march only
    # competes with Adam for the same SMs -- so it is off unless asked for
    lookahead = os.environ.get("LNRF_LOOKAHEAD", "0") == "1"
    gstep, graph_note = None, ("cuda graph (one capture per sample-buffer size)" +
                               ("; look-ahead: the parameter-independent near/far + march of batch k runs on a second stream beside "
                                "the network/backward/Adam of batch k-1 -- every step still marches one batch and trains on one" if lookahead else ""))
    if os.environ.get("LNRF_NO_GRAPH", "0") != "1":
        try:
            gstep = GraphedTrainStep(step, N_RAYS, lookahead=lookahead)
            gstep.capture(*dev_batches[0])
            for i in range(3):
                gstep(*dev_batches[i % n_batches])
        except Exception as e:  # never silently: the JSON line says which path was timed
            gstep, graph_note = None, f"eager (graph capture failed: {type(e).__name__}: {e})"[:300]
            torch.cuda.synchronize()
    run = gstep if gstep is not None else (lambda *b: step(*b))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = _native.launch_count()
    ms = timed_loop(lambda i: run(*dev_batches[i % n_batches]), args.steps)
    launches = _native.launch_count() - launches0
    if gstep is not None:  # replays do not pass through the C ABI: count the launches the captured step contains
        launches = eager_launches  # same step, same number of steps, counted when it was issued through the C ABI
    clk = clocks.stop() if rank == 0 else None
    in_sync = None
    if world > 1 and step.fused_optimizer:  # every rank must hold the same fp16 table after the sharded optimizer steps
        chk = model.encoder._shadow_f16.float().abs().sum().double().reshape(1)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        in_sync = bool(lo_.item() == hi_.item())
    seq_ms = None
    if gstep is not None and lookahead:  # the same graph without the cross-step overlap, for reference
        try:
            model.update_mean_count()
            g2 = GraphedTrainStep(step, N_RAYS, lookahead=False)
            g2.capture(*dev_batches[0])
            for i in range(3):
                g2(*dev_batches[i % n_batches])
            seq_ms = timed_loop(lambda i: g2(*dev_batches[i % n_batches]), args.steps) / args.steps
            del g2
        except Exception as e:
            seq_ms = f"failed: {type(e).__name__}: {e}"[:200]
    ms_per_step = ms / args.steps
    value = world * N_RAYS * args.steps / (ms * 1e-3)

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H loss, every step ----------------------------------------
    h2d = sum(x.numel() * x.element_size() for x in host_batches[0])

    def e2e_iter(i):
        hb = host_batches[i % n_batches]
        if gstep is not None:
            loss, _ = gstep(*hb)  # static device buffers are filled straight from pinned memory (non_blocking copies)
        else:
            loss, _ = step(*(x.to(dev, non_blocking=True) for x in hb))
        float(loss.item())

    e2e_ms = timed_loop(e2e_iter, args.steps)
    e2e_value = world * N_RAYS * args.steps / (e2e_ms * 1e-3)

    # ---- per-kernel roofline (events recorded live inside the timed region above) -----------------------------------
    peaks = measured_peaks()
    n_params = sum(p.numel() for p in model.parameters())
    m_pad = int(statistics.mean(points))
    kern = timed.summary()
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    table = {}
    for name, k in kern.items():
        a = ALGO[name]
        calls_per_step = k["calls"] / args.steps
        if a["bound"] == "hbm":
            byts = (a.get("per_ray", 0) * N_RAYS + a.get("per_sample", 0) * actual + a.get("per_sample_padded", 0) * m_pad +
                    a.get("per_param", 0) * (n_params // world))  # sharded optimizer: each rank updates 1/world of the table
            ach = byts / (k["mean_ms"] * 1e-3) / 1e9
            table[name] = dict(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"], mean_ms=k["mean_ms"],
                               calls_per_step=calls_per_step, algorithmic_bytes=byts, traffic=traffic.get(name))
        else:
            fl = a["flops_per_sample_padded"] * m_pad
            ach = fl / (k["mean_ms"] * 1e-3) / 1e12
            table[name] = dict(bound="tensor", achieved=ach, peak=peaks["tensor"], unit="TFLOP/s", frac=ach / peaks["tensor"],
                               mean_ms=k["mean_ms"], calls_per_step=calls_per_step, algorithmic_flops=fl, traffic=traffic.get(name))
    dominant = max(table, key=lambda n: kern[n]["total_ms"]) if table else None
    roofline = None
    if dominant:
        d = table[dominant]
        roofline = dict(bound=d["bound"], achieved=d["achieved"], peak=d["peak"], unit=d["unit"], frac=d["frac"], traffic=d["traffic"],
                        kernel=dominant, mean_ms=d["mean_ms"], peak_source=peaks["source"] + (" (sustained)" if d["bound"] == "tensor" else ""),
                        share_of_step=kern[dominant]["total_ms"] / sum(k["total_ms"] for k in kern.values()),
                        timing="CUDA events around each C-ABI launch during the eager pass of the same step (graph replays cannot be bracketed)")

    # ---- render: one full 800x800 view, tile-sharded over ranks (no collective but the final gather) -------------------
    # (at N = 1 the rays are in plain row-major order, as the reference renders them)
    render = None
    if not args.no_render:
        ro, rd, _ = get_rays_np(sc.poses[0], sc.intrinsics, sc.H, sc.W)
        # N > 1: 32 x 32 pixel tiles dealt round-robin over the ranks (contiguous row ranges leave the object to the middle ranks)
        mine = tile_shard_indices(sc.H, sc.W, rank, world).numpy() if world > 1 else np.arange(sc.H * sc.W)
        ro_d, rd_d = torch.from_numpy(ro[mine]).to(dev), torch.from_numpy(rd[mine]).to(dev)
        model.eval()
        model.render_schedule = os.environ.get("LNRF_RENDER_SCHEDULE", "fast")  # "reference": run_cuda's n_step rule, bit for bit
        # real samples of this rank's rays (the slots of a round include the zero padding the API mandates, and the fast schedule
        # pads more): counted once, outside the timed region, by the one-shot marcher on the same rays without a sample budget
        from laenerf_b200 import raymarching as _rm
        real = 0
        with torch.no_grad():
            for c0 in range(0, ro_d.shape[0], 65536):
                o_, d_ = ro_d[c0:c0 + 65536].contiguous(), rd_d[c0:c0 + 65536].contiguous()
                ne_, fa_ = _rm.near_far_from_aabb(o_, d_, model.aabb_infer, model.min_near)
                cnt_ = torch.zeros(2, dtype=torch.int32, device=dev)
                _rm.march_rays_train(o_, d_, model.bound, model.density_bitfield, model.cascade, model.grid_size, ne_, fa_, cnt_, -1, False,
                                     128, True, 0, 1024)
                real += int(cnt_[0].item())
        frames, samples = 0, 0
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e0, e1 = ev(), ev()
        marks = []
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for _ in range(2):  # warm-up frames, final gather included (the first collective of a shape pays NCCL's lazy set-up)
                gather_tiles(model.render(ro_d, rd_d, perturb=False, bg_color=1)["image"], sc.H, sc.W, rank, world)
            barrier()
            e0.record()
            for _ in range(3):
                a, b, c = ev(), ev(), ev()
                a.record()
                out = model.render(ro_d, rd_d, perturb=False, bg_color=1)
                b.record()
                img = gather_tiles(out["image"], sc.H, sc.W, rank, world)
                c.record()
                marks.append((a, b, c))
                samples += out["num_points"]
                frames += 1
            e1.record()
            barrier()
        loop_ms = sum(a.elapsed_time(b) for a, b, _ in marks) / frames
        gather_ms = sum(b.elapsed_time(c) for _, b, c in marks) / frames
        tt = torch.tensor([e0.elapsed_time(e1), float(samples), loop_ms, gather_ms, float(real)], device=dev, dtype=torch.float64)
        if world > 1:
            tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            tt = torch.stack([tmax[0], tsum[1], tmax[2], tmax[3], tsum[4]])
        rms, rs, real_all = float(tt[0].item()), float(tt[1].item()), float(tt[4].item())
        render = dict(value=real_all * frames / (rms * 1e-3) / 1e6, unit="Msamples/s", ms_per_frame=rms / frames, rays_per_frame=int(ro.shape[0]),
                      samples_per_frame=real_all, sample_slots_per_frame=rs / frames, slots_msamples_per_s=rs / (rms * 1e-3) / 1e6,
                      note="value counts REAL samples (occupied-cell samples of the frame's rays); slots include the zero padding of every round",
                      rays_per_s=ro.shape[0] * frames / (rms * 1e-3), image_shape=list(img.shape),
                      rounds=out.get("rounds"), schedule=model.render_schedule, render_loop_ms=float(tt[2].item()), gather_ms=float(tt[3].item()),
                      sharding=("32x32-pixel tiles dealt round-robin over the ranks; all_gather + scatter by the known index lists"
                                if world > 1 else "single GPU, row-major rays"))
        model.train()

    if rank != 0:
        _finish(world)
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import cpu_renderer
        n = int(os.environ.get("LNRF_CPU_RAYS", "1024"))
        hb = host_batches[0]
        sec, threads, used = cpu_renderer.time_train_steps(hb[0][:n].clone(), hb[1][:n].clone(), hb[2][:n].clone(), steps=2, warmup=1,
                                                           num_steps=512, bound=sc.bound, min_near=sc.min_near)
        cpu = dict(value=used / sec, unit="rays/s", cores=threads, kind="port",
                   sample=f"{used} rays x 512 uniform samples/ray per step (non-cuda_ray renderer restated in oracle/cpu_renderer.py), 2 timed + 1 warm-up steps")

    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "samples_per_step_padded": m_pad, "samples_per_step": actual,
                   "samples_per_ray": actual / N_RAYS, "scene_occupancy": sc.occupancy_fraction(),
                   "l2": "no explicit flush: one step touches ~245 MB (fp32 table + grads + Adam moments + fp16 copies) > 126 MB L2",
                   "occupancy_update": "excluded (row f-2 of SURVEY.md section 8: fixed procedural occupancy grid)",
                   "parallelism": (f"ray-sharded dp{world}: reduce-scatter of the fp16 hash-grid gradient, Adam on a 1/{world} table slice per "
                                   f"rank, all-gather of the fp16 table (ZeRO-1 style)") if world > 1 else "single GPU"},
        "clocks": clk,
        "step_mode": graph_note,
        "graph_sequential_ms_per_step": seq_ms,
        "replicas_in_sync": in_sync,
        "exchange_sync": (None if world == 1 or not step.fused_optimizer or step.optimizer.p2p is None else
                          ("inside the kernels (signal + poll on peer-mapped flag words)" if step.optimizer.inkernel_sync else
                           "two symmetric-memory barrier launches around the kernel")),
        "exchange_timing_us": (step.optimizer.exchange_timing() if world > 1 and step.fused_optimizer and os.environ.get("LNRF_TIME_EXCHANGE") == "1"
                               else None),
        "exchange": (None if world == 1 or not step.fused_optimizer else
                     ("one fused kernel over NVLink peer memory (torch symmetric memory): average of the ranks' fp16 gradients + Adam on "
                      "a 1/N slice + store of the new fp16 values into every rank's table" if step.optimizer.p2p is not None else
                      "NCCL reduce-scatter + Adam on a 1/N slice + NCCL all-gather (symmetric memory unavailable: " +
                      str(getattr(step.optimizer, "p2p_error", "disabled")) + ")")),
        "eager": {"ms_per_step": eager_ms / args.steps, "value": world * N_RAYS * args.steps / (eager_ms * 1e-3), "unit": "rays/s",
                  "note": "same step issued launch by launch from Python (the drop-in modules without graph capture)"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": table,
        "render": render,
        "cpu_baseline": cpu,
        "train_msamples_per_s": world * actual * args.steps / (ms * 1e-3) / 1e6,
    }
    print(json.dumps(line), file=_OUT, flush=True)
    _finish(world)


def _finish(world):
    """Multi-rank runs leave without tearing NCCL down: destroy_process_group() with captured collectives still alive in
    CUDA graphs was observed to hang until the launcher's timeout (2-GPU run of round 1).  All collectives are complete
    (every rank passed the final barrier + synchronize), so the processes simply exit."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
