"""Drop-in for the reference's `raymarching` package (raymarching/raymarching.py:19-461) on liblaenerf_b200.so.

Same public names, positional signatures, return shapes and AMP contract (`custom_fwd(cast_inputs=float32)`);
the native back-end is the C-ABI library (no pybind11, no torch headers).  Differences that are NOT observable
through the API, all on purpose:
  * the sample buffers are `torch.empty` -- the kernels zero-fill every row a ray does not write themselves, so the
    caller sees exactly what `torch.zeros` + the reference kernel leave behind, minus three memset launches;
  * `rays` rows are in ray-id order with prefix-sum offsets (the reference's order depends on atomic arrival);
  * `composite_rays_train.backward` lets the kernel clear the gradient rows no ray covers.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from . import _native as N

__all__ = [
    "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
    "composite_rays_train", "march_rays", "march_rays_distill", "composite_rays", "composite_rays_distill",
    "compact_alive", "composite_loss_train",
]

_scratch = {}  # (device index, stream, kind) -> zero-initialised int64 scratch the kernels keep zeroed
_DEBUG_LAYOUT = os.environ.get("LNRF_DEBUG_LAYOUT", "0") == "1"  # composite_loss_train: verify the canonical `rays` layout (one sync)


def _get_scratch(kind: str, nbytes: int, device) -> torch.Tensor:
    key = (device.index, N.stream(), kind)
    t = _scratch.get(key)
    if t is None or t.numel() * 8 < nbytes:
        if len(_scratch) >= _MAX_SCRATCH:  # oldest first (dicts keep insertion order)
            del _scratch[next(iter(_scratch))]
        t = torch.zeros(max(1024, (nbytes + 7) // 8 * 2), dtype=torch.int64, device=device)
        _scratch[key] = t
    return t


def _cuda(t):
    return t if t.is_cuda else t.cuda()


_MAX_SCRATCH = 32  # cache entries (streams x kinds); warm-up side streams and re-captures must not grow it without bound


def _scratch_reset(device) -> None:
    """After a failed launch the look-back status words may be left non-zero and every later launch would read stale flags:
    drop the cached scratch of the device so the next call starts from fresh zeros."""
    for k in [k for k in _scratch if k[0] == device.index]:
        del _scratch[k]


def _check(status: int, device) -> None:
    if status != 0:
        _scratch_reset(device)
        N.check(status)


def _in(t, dtype, name):
    """Read-only kernel input: the reference raises through CHECK_CUDA / CHECK_CONTIGUOUS and dispatches on the dtype
    (raymarching.cu: AT_DISPATCH_FLOATING_TYPES_AND_HALF); here the tensor must be on the GPU and is brought to the one dtype the
    kernels are built for (fp32 values / int32 indices / uint8 bitfields) -- never reinterpreted."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        if dtype is torch.float32 and t.dtype in (torch.float16, torch.bfloat16, torch.float64):
            t = t.float()
        elif dtype is torch.uint8 and t.dtype is torch.bool:
            t = t.view(torch.uint8)
        else:
            raise RuntimeError(f"{name} must be {dtype}, but got {t.dtype}")
    return t.contiguous()


def _inout(t, dtype, name):
    """Tensor the kernel updates in place: it must already be what the kernel writes (a converted copy would lose the result)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, but got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    return t


class _on:
    """Launch on the device the tensors live on (the reference has no device guard; kernels would launch on the current device)."""

    def __init__(self, t):
        self.guard = None if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *a):
        if self.guard is not None:
            self.guard.__exit__(*a)


class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        aabb = _cuda(aabb).contiguous()
        n = rays_o.shape[0]
        nears = torch.empty(n, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(n, dtype=rays_o.dtype, device=rays_o.device)
        N.check(N.lib().lnrf_near_far_from_aabb(N.ptr(rays_o), N.ptr(rays_d), N.ptr(aabb), n, float(min_near), N.ptr(nears),
                                               N.ptr(fars), N.stream()))
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        n = rays_o.shape[0]
        coords = torch.empty(n, 2, dtype=rays_o.dtype, device=rays_o.device)
        N.check(N.lib().lnrf_sph_from_ray(N.ptr(rays_o), N.ptr(rays_d), float(radius), n, N.ptr(coords), N.stream()))
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords).int().contiguous()
        n = coords.shape[0]
        indices = torch.empty(n, dtype=torch.int32, device=coords.device)
        N.check(N.lib().lnrf_morton3D(N.ptr(coords), n, N.ptr(indices), N.stream()))
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices).int().contiguous()
        n = indices.shape[0]
        coords = torch.empty(n, 3, dtype=torch.int32, device=indices.device)
        N.check(N.lib().lnrf_morton3D_invert(N.ptr(indices), n, N.ptr(coords), N.stream()))
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _cuda(grid).contiguous()
        n = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(n, dtype=torch.uint8, device=grid.device)
        else:
            ctx.mark_dirty(bitfield)  # written in place through a raw pointer: bump torch's version counter (NeRFNetwork.occupied_box keys on it)
        N.check(N.lib().lnrf_packbits(N.ptr(grid), n, float(thresh), N.ptr(bitfield), N.stream()))
        return bitfield


packbits = _packbits.apply


class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, out=None, occupied_box=None):
        """raymarching.py:176-240.  `out` (not in the reference): (xyzs, dirs, deltas, rays) to march into instead of new tensors --
        a software-pipelined training loop alternates between two sample buffers (nerf.GraphedTrainStep).  `occupied_box` (not in the
        reference): device float[6] around every occupied cell (nerf.NeRFNetwork.occupied_box); rays end where they leave it, with
        identical samples / counts / offsets."""
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        density_bitfield = _in(_cuda(density_bitfield), torch.uint8, "density_bitfield")
        nears, fars = _in(nears, torch.float32, "nears"), _in(fars, torch.float32, "fars")
        dev = rays_o.device
        n = rays_o.shape[0]
        M = n * max_steps
        if not force_all_rays and mean_count > 0:  # raymarching.py:199-203
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        if out is not None:
            xyzs, dirs, deltas, rays = out
            for t_, shp, dt_, nm in ((xyzs, (M, 3), rays_o.dtype, "xyzs"), (dirs, (M, 3), rays_o.dtype, "dirs"),
                                     (deltas, (M, 2), rays_o.dtype, "deltas"), (rays, (n, 3), torch.int32, "rays")):
                if tuple(t_.shape) != shp or t_.dtype != dt_ or t_.device != dev or not t_.is_contiguous():
                    raise RuntimeError(f"march_rays_train: out.{nm} must be a contiguous {dt_} tensor of shape {shp} on {dev}")
        else:
            xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
            dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
            deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
            rays = torch.empty(n, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        step_counter = _inout(step_counter, torch.int32, "step_counter")
        noises = torch.rand(n, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(n, dtype=rays_o.dtype, device=dev)
        lib = N.lib()
        nbytes = lib.lnrf_march_rays_train_scratch_bytes(n)
        scratch = _get_scratch("march", nbytes, dev)
        with _on(rays_o):
            if occupied_box is not None:
                occupied_box = _in(occupied_box, torch.float32, "occupied_box")
            _check(lib.lnrf_march_rays_train_clipped(N.ptr(rays_o), N.ptr(rays_d), N.ptr(density_bitfield), float(bound), float(dt_gamma),
                                                     int(max_steps), n, int(C), int(H), M, N.ptr(nears), N.ptr(fars), N.ptr(xyzs),
                                                     N.ptr(dirs), N.ptr(deltas), N.ptr(rays), N.ptr(step_counter), N.ptr(noises),
                                                     N.ptr(occupied_box), N.ptr(scratch), scratch.numel() * 8, N.stream()), dev)
        if force_all_rays or mean_count <= 0:  # raymarching.py:222-231 (first epochs only)
            m = int(step_counter[0].item())
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, deltas = _in(sigmas, torch.float32, "sigmas"), _in(rgbs, torch.float32, "rgbs"), _in(deltas, torch.float32, "deltas")
        rays = _in(rays, torch.int32, "rays")
        M, n = sigmas.shape[0], rays.shape[0]
        weights_sum = torch.empty(n, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(n, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(n, 3, dtype=sigmas.dtype, device=sigmas.device)
        N.check(N.lib().lnrf_composite_rays_train_forward(N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays), M, n,
                                                          float(T_thresh), N.ptr(weights_sum), N.ptr(depth), N.ptr(image),
                                                          N.stream()))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, n, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):  # grad_depth is dropped (raymarching.py:275)
        grad_weights_sum, grad_image = _in(grad_weights_sum, torch.float32, "grad_weights_sum"), _in(grad_image, torch.float32, "grad_image")
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, n, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        N.check(N.lib().lnrf_composite_rays_train_backward(N.ptr(grad_weights_sum), N.ptr(grad_image), N.ptr(sigmas), N.ptr(rgbs),
                                                           N.ptr(deltas), N.ptr(rays), N.ptr(weights_sum), N.ptr(image), M, n,
                                                           float(T_thresh), N.ptr(grad_sigmas), N.ptr(grad_rgbs), 0, N.stream()))
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


class _composite_loss_train(Function):
    """Row f-5 (fused tier, no reference twin): composite_rays_train + background blend + depth normalisation + the trainer's
    MSE loss in ONE launch, and the whole backward of that tail in one more (include/laenerf_b200.h).  What it replaces:
    renderer.py:324-329 and nerf/utils.py:592,633 -- about a dozen elementwise / reduction launches each way.
    Returns (loss, weights_sum, depth, image); only `loss` is differentiable (w.r.t. sigmas and rgbs).

    grad_scale: the device scalar that will come back as dL/dloss (the AMP loss scale the caller hands to autograd.backward).  When
    given, forward and backward are ONE launch (lnrf_composite_loss_train_forward_backward): the gradients are formed while the
    ray's samples are still in L1 and the backward only hands them over.  If a different gradient arrives after all, the backward
    falls back to its own launch -- the results are the same bits either way."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, gt_rgb, bg_color, nears, fars, T_thresh=1e-4, grad_scale=None):
        sigmas, rgbs, deltas = _in(sigmas, torch.float32, "sigmas"), _in(rgbs, torch.float32, "rgbs"), _in(deltas, torch.float32, "deltas")
        rays = _in(rays, torch.int32, "rays")
        M, n = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        if _DEBUG_LAYOUT and n > 1:
            # the backward's zero-fill assumes march_rays_train's canonical layout: rows in ray-id order, offsets = exclusive prefix
            # sum of the counts (a `rays` tensor in the reference's atomic-arrival order must go through composite_rays_train)
            kept = rays[:, 1].long() + rays[:, 2].long() <= M
            off_ok = rays[1:, 1] == rays[:-1, 1] + rays[:-1, 2]
            if not (bool((rays[:, 0] == torch.arange(n, device=dev, dtype=torch.int32)).all()) and bool(off_ok[kept[1:] & kept[:-1]].all())):
                raise RuntimeError("composite_loss_train: `rays` is not in march_rays_train's canonical (ray-id ordered, prefix-sum) layout")
        gt_rgb = _cuda(gt_rgb).contiguous().view(n, 3)
        if torch.is_tensor(bg_color) and bg_color.numel() == 3 * n:
            bg, bg_scalar = _cuda(bg_color).contiguous().view(n, 3), 0.0
        elif torch.is_tensor(bg_color):
            if bg_color.numel() != 1:
                raise RuntimeError("composite_loss_train: bg_color must be a scalar or one colour per ray")
            bg, bg_scalar = None, float(bg_color)
        else:
            bg, bg_scalar = None, float(bg_color)
        weights_sum = torch.empty(n, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        image_raw = torch.empty(n, 3, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lib = N.lib()
        nbytes = lib.lnrf_composite_loss_scratch_bytes(n)
        scratch = _get_scratch("composite_loss", nbytes, dev)
        ctx.ready = None
        fuse = (grad_scale is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and
                grad_scale.is_cuda and grad_scale.dtype == torch.float32 and grad_scale.numel() == 1)
        if fuse:
            grad_sigmas = torch.empty_like(sigmas)  # the kernel writes every element (zero_fill contract)
            grad_rgbs = torch.empty_like(rgbs)
            N.check(lib.lnrf_composite_loss_train_forward_backward(
                N.ptr(grad_scale), N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays), N.ptr(gt_rgb), N.ptr(bg), bg_scalar, N.ptr(nears),
                N.ptr(fars), M, n, float(T_thresh), N.ptr(weights_sum), N.ptr(depth), N.ptr(image), N.ptr(image_raw), N.ptr(loss),
                N.ptr(grad_sigmas), N.ptr(grad_rgbs), N.ptr(scratch), nbytes, N.stream()))
            ctx.ready = (grad_scale.data_ptr(), grad_scale._version, grad_sigmas, grad_rgbs)
        else:
            N.check(lib.lnrf_composite_loss_train_forward(N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays), N.ptr(gt_rgb), N.ptr(bg),
                                                          bg_scalar, N.ptr(nears), N.ptr(fars), M, n, float(T_thresh), N.ptr(weights_sum),
                                                          N.ptr(depth), N.ptr(image), N.ptr(image_raw), N.ptr(loss), N.ptr(scratch),
                                                          nbytes, N.stream()))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, gt_rgb, bg, weights_sum, image, image_raw)
        ctx.dims = [M, n, T_thresh, bg_scalar]
        ctx.mark_non_differentiable(weights_sum, depth, image)
        ctx.set_materialize_grads(False)  # no zero tensors (three fill launches per step) for the outputs nobody differentiates
        return loss, weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_loss, *_):
        if grad_loss is None:
            return (None,) * 10
        if ctx.ready is not None and grad_loss.data_ptr() == ctx.ready[0] and grad_loss._version == ctx.ready[1]:
            return ctx.ready[2], ctx.ready[3], None, None, None, None, None, None, None, None  # formed by the forward's launch
        sigmas, rgbs, deltas, rays, gt_rgb, bg, weights_sum, image, image_raw = ctx.saved_tensors
        M, n, T_thresh, bg_scalar = ctx.dims
        grad_loss = grad_loss.to(device=sigmas.device, dtype=torch.float32).contiguous()
        grad_sigmas = torch.empty_like(sigmas)  # the kernel writes every element (zero_fill contract)
        grad_rgbs = torch.empty_like(rgbs)
        N.check(N.lib().lnrf_composite_loss_train_backward(N.ptr(grad_loss), N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays),
                                                           N.ptr(gt_rgb), N.ptr(bg), bg_scalar, N.ptr(weights_sum), N.ptr(image),
                                                           N.ptr(image_raw), M, n, float(T_thresh), N.ptr(grad_sigmas),
                                                           N.ptr(grad_rgbs), N.stream()))
        return grad_sigmas, grad_rgbs, None, None, None, None, None, None, None, None


composite_loss_train = _composite_loss_train.apply


def _march_infer(distill, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, edit_bitfield, C, H,
                 near, far, align, perturb, dt_gamma, max_steps):
    rays_o = _cuda(rays_o).contiguous().view(-1, 3)
    rays_d = _cuda(rays_d).contiguous().view(-1, 3)
    rays_alive, rays_t = _in(rays_alive, torch.int32, "rays_alive"), _in(rays_t, torch.float32, "rays_t")
    near, far = _in(near, torch.float32, "near"), _in(far, torch.float32, "far")
    density_bitfield = _in(density_bitfield, torch.uint8, "density_bitfield")
    edit_bitfield = _in(edit_bitfield, torch.uint8, "edit_bitfield")
    dev = rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)  # raymarching.py:331-332: always pads, a full `align` when already aligned
    xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
    deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
    noises = torch.rand(n_alive, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(n_alive, dtype=rays_o.dtype, device=dev)
    lib = N.lib()
    if distill:
        edit_occ = torch.empty(M, dtype=torch.bool, device=dev)
        N.check(lib.lnrf_march_rays_distill(n_alive, n_step, N.ptr(rays_alive), N.ptr(rays_t), N.ptr(rays_o), N.ptr(rays_d),
                                            float(bound), float(dt_gamma), int(max_steps), int(C), int(H), N.ptr(density_bitfield),
                                            N.ptr(edit_bitfield), N.ptr(near), N.ptr(far), N.ptr(xyzs), N.ptr(dirs), N.ptr(deltas),
                                            N.ptr(edit_occ), N.ptr(noises), M, N.stream()))
        return xyzs, dirs, deltas, edit_occ
    N.check(lib.lnrf_march_rays(n_alive, n_step, N.ptr(rays_alive), N.ptr(rays_t), N.ptr(rays_o), N.ptr(rays_d), float(bound),
                                float(dt_gamma), int(max_steps), int(C), int(H), N.ptr(density_bitfield), N.ptr(near), N.ptr(far),
                                N.ptr(xyzs), N.ptr(dirs), N.ptr(deltas), N.ptr(noises), M, N.stream()))
    return xyzs, dirs, deltas


class _march_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        return _march_infer(False, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, None, C, H, near,
                            far, align, perturb, dt_gamma, max_steps)


march_rays = _march_rays.apply


class _march_rays_distill(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, edit_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        return _march_infer(True, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, edit_bitfield, C, H,
                            near, far, align, perturb, dt_gamma, max_steps)


march_rays_distill = _march_rays_distill.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        sigmas, rgbs, deltas = _in(sigmas, torch.float32, "sigmas"), _in(rgbs, torch.float32, "rgbs"), _in(deltas, torch.float32, "deltas")
        rays_alive, rays_t = _inout(rays_alive, torch.int32, "rays_alive"), _inout(rays_t, torch.float32, "rays_t")
        weights_sum, depth, image = (_inout(weights_sum, torch.float32, "weights_sum"), _inout(depth, torch.float32, "depth"),
                                     _inout(image, torch.float32, "image"))
        with _on(sigmas):
            N.check(N.lib().lnrf_composite_rays(n_alive, n_step, float(T_thresh), N.ptr(rays_alive), N.ptr(rays_t), N.ptr(sigmas),
                                                N.ptr(rgbs), N.ptr(deltas), N.ptr(weights_sum), N.ptr(depth), N.ptr(image), N.stream()))
        return tuple()


composite_rays = _composite_rays.apply


class _composite_rays_distill(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, weights_edit_sum, depth, depth_edit,
                image, int_edit, T_thresh=1e-2):
        sigmas, rgbs, deltas = _in(sigmas, torch.float32, "sigmas"), _in(rgbs, torch.float32, "rgbs"), _in(deltas, torch.float32, "deltas")
        rays_alive, rays_t = _inout(rays_alive, torch.int32, "rays_alive"), _inout(rays_t, torch.float32, "rays_t")
        weights_sum, depth, image = (_inout(weights_sum, torch.float32, "weights_sum"), _inout(depth, torch.float32, "depth"),
                                     _inout(image, torch.float32, "image"))
        weights_edit_sum, depth_edit = _inout(weights_edit_sum, torch.float32, "weights_edit_sum"), _inout(depth_edit, torch.float32, "depth_edit")
        int_edit = _in(int_edit, torch.uint8, "int_edit")
        with _on(sigmas):
            N.check(N.lib().lnrf_composite_rays_distill(n_alive, n_step, float(T_thresh), N.ptr(rays_alive), N.ptr(rays_t),
                                                        N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(weights_sum),
                                                        N.ptr(weights_edit_sum), N.ptr(depth), N.ptr(depth_edit), N.ptr(int_edit),
                                                        N.ptr(image), N.stream()))
        return tuple()


composite_rays_distill = _composite_rays_distill.apply


def compact_alive(rays_alive: torch.Tensor, n_alive: int, out: torch.Tensor | None = None, count: torch.Tensor | None = None):
    """Device-side `rays_alive[rays_alive >= 0]` (renderer.py:375) without the boolean-mask kernels: returns
    (out, count) where out[:count] holds the surviving ray ids in order and count is a device int32[1]."""
    dev = rays_alive.device
    if out is None:
        out = torch.empty(max(n_alive, 1), dtype=torch.int32, device=dev)
    if count is None:
        count = torch.empty(1, dtype=torch.int32, device=dev)
    lib = N.lib()
    scratch = _get_scratch("compact", lib.lnrf_compact_alive_scratch_bytes(n_alive), dev)
    _check(lib.lnrf_compact_alive(N.ptr(rays_alive), n_alive, N.ptr(out), N.ptr(count), N.ptr(scratch), scratch.numel() * 8,
                                  N.stream()), dev)
    return out, count
