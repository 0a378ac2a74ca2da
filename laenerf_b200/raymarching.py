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


def _get_scratch(kind: str, nbytes: int, device) -> torch.Tensor:
    key = (device.index, N.stream(), kind)
    t = _scratch.get(key)
    if t is None or t.numel() * 8 < nbytes:
        t = torch.zeros(max(1024, (nbytes + 7) // 8 * 2), dtype=torch.int64, device=device)
        _scratch[key] = t
    return t


def _cuda(t):
    return t if t.is_cuda else t.cuda()


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
        N.check(N.lib().lnrf_packbits(N.ptr(grid), n, float(thresh), N.ptr(bitfield), N.stream()))
        return bitfield


packbits = _packbits.apply


class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        density_bitfield = _cuda(density_bitfield).contiguous()
        dev = rays_o.device
        n = rays_o.shape[0]
        M = n * max_steps
        if not force_all_rays and mean_count > 0:  # raymarching.py:199-203
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
        rays = torch.empty(n, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(n, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(n, dtype=rays_o.dtype, device=dev)
        lib = N.lib()
        nbytes = lib.lnrf_march_rays_train_scratch_bytes(n)
        scratch = _get_scratch("march", nbytes, dev)
        N.check(lib.lnrf_march_rays_train(N.ptr(rays_o), N.ptr(rays_d), N.ptr(density_bitfield), float(bound), float(dt_gamma),
                                          int(max_steps), n, int(C), int(H), M, N.ptr(nears), N.ptr(fars), N.ptr(xyzs),
                                          N.ptr(dirs), N.ptr(deltas), N.ptr(rays), N.ptr(step_counter), N.ptr(noises),
                                          N.ptr(scratch), scratch.numel() * 8, N.stream()))
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
        sigmas, rgbs, deltas = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous()
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
        grad_weights_sum, grad_image = grad_weights_sum.contiguous(), grad_image.contiguous()
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
    Returns (loss, weights_sum, depth, image); only `loss` is differentiable (w.r.t. sigmas and rgbs)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, gt_rgb, bg_color, nears, fars, T_thresh=1e-4):
        sigmas, rgbs, deltas = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous()
        M, n = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
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
            return (None,) * 9
        sigmas, rgbs, deltas, rays, gt_rgb, bg, weights_sum, image, image_raw = ctx.saved_tensors
        M, n, T_thresh, bg_scalar = ctx.dims
        grad_loss = grad_loss.to(device=sigmas.device, dtype=torch.float32).contiguous()
        grad_sigmas = torch.empty_like(sigmas)  # the kernel writes every element (zero_fill contract)
        grad_rgbs = torch.empty_like(rgbs)
        N.check(N.lib().lnrf_composite_loss_train_backward(N.ptr(grad_loss), N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays),
                                                           N.ptr(gt_rgb), N.ptr(bg), bg_scalar, N.ptr(weights_sum), N.ptr(image),
                                                           N.ptr(image_raw), M, n, float(T_thresh), N.ptr(grad_sigmas),
                                                           N.ptr(grad_rgbs), N.stream()))
        return grad_sigmas, grad_rgbs, None, None, None, None, None, None, None


composite_loss_train = _composite_loss_train.apply


def _march_infer(distill, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, edit_bitfield, C, H,
                 near, far, align, perturb, dt_gamma, max_steps):
    rays_o = _cuda(rays_o).contiguous().view(-1, 3)
    rays_d = _cuda(rays_d).contiguous().view(-1, 3)
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
        sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
        N.check(N.lib().lnrf_composite_rays(n_alive, n_step, float(T_thresh), N.ptr(rays_alive), N.ptr(rays_t), N.ptr(sigmas),
                                            N.ptr(rgbs), N.ptr(deltas), N.ptr(weights_sum), N.ptr(depth), N.ptr(image), N.stream()))
        return tuple()


composite_rays = _composite_rays.apply


class _composite_rays_distill(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, weights_edit_sum, depth, depth_edit,
                image, int_edit, T_thresh=1e-2):
        sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
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
    N.check(lib.lnrf_compact_alive(N.ptr(rays_alive), n_alive, N.ptr(out), N.ptr(count), N.ptr(scratch), scratch.numel() * 8,
                                   N.stream()))
    return out, count
