"""ctypes bindings of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
package (laenerf_b200) never does.  All functions take and return numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblaenerf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


u32, f32c, i32c = C.c_uint32, C.c_float, C.c_int32


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32c(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(rays_o), _p(rays_d), f32c(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords).reshape(-1, 3)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_p(coords), u32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices).reshape(-1)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), u32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid).reshape(-1)
    N = grid.shape[0] // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), u32(N), f32c(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C_, H, M, nears, fars, noises, counter=None):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    bitfield, nears, fars, noises = _u8(bitfield), _f32(nears), _f32(fars), _f32(noises)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32) if counter is None else _i32(counter).copy()
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32c(bound), f32c(dt_gamma), u32(max_steps), u32(N),
                               u32(C_), u32(H), u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays),
                               _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), u32(M), u32(N), f32c(T_thresh),
                                           _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    grad_ws, grad_image, ws, image = _f32(grad_ws), _f32(grad_image), _f32(ws), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays), _p(ws),
                                            _p(image), u32(M), u32(N), f32c(T_thresh), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C_, H, nears, fars, noises, M_rows,
               dt_gamma=0.0, max_steps=1024, edit_bitfield=None):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t, nears, fars, noises = _i32(rays_alive), _f32(rays_t), _f32(nears), _f32(fars), _f32(noises)
    bitfield = _u8(bitfield)
    xyzs, dirs, deltas = np.zeros((M_rows, 3), np.float32), np.zeros((M_rows, 3), np.float32), np.zeros((M_rows, 2), np.float32)
    edit_occ = None
    if edit_bitfield is not None:
        edit_bitfield = _u8(edit_bitfield)
        edit_occ = np.zeros(M_rows, np.uint8)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), f32c(bound),
                         f32c(dt_gamma), u32(max_steps), u32(C_), u32(H), _p(bitfield), _p(edit_bitfield), _p(nears),
                         _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(edit_occ), _p(noises))
    return (xyzs, dirs, deltas) if edit_occ is None else (xyzs, dirs, deltas, edit_occ)


def march_rays_track(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C_, H, nears, fars, noises, M_rows,
                     dt_gamma=0.0, max_steps=1024):
    """march_rays plus, per alive slot, whether a delta emitted in this round was not an exact difference (orc_march_rays_track; test
    bookkeeping).  Returns (xyzs, dirs, deltas, inexact)."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t, nears, fars, noises = _i32(rays_alive), _f32(rays_t), _f32(nears), _f32(fars), _f32(noises)
    bitfield = _u8(bitfield)
    xyzs, dirs, deltas = np.zeros((M_rows, 3), np.float32), np.zeros((M_rows, 3), np.float32), np.zeros((M_rows, 2), np.float32)
    inexact = np.zeros(max(int(n_alive), 1), np.uint8)
    lib().orc_march_rays_track(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), f32c(bound), f32c(dt_gamma),
                               u32(max_steps), u32(C_), u32(H), _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas),
                               _p(noises), _p(inexact))
    return xyzs, dirs, deltas, inexact[:n_alive]


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2,
                   weights_edit_sum=None, depth_edit=None, edit_occ=None):
    """Returns updated copies: (rays_alive, rays_t, weights_sum, depth, image[, weights_edit_sum, depth_edit])."""
    rays_alive, rays_t = _i32(rays_alive).copy(), _f32(rays_t).copy()
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    weights_sum, depth, image = _f32(weights_sum).copy(), _f32(depth).copy(), _f32(image).copy()
    distill = edit_occ is not None
    if distill:
        weights_edit_sum, depth_edit, edit_occ = _f32(weights_edit_sum).copy(), _f32(depth_edit).copy(), _u8(edit_occ)
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32c(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs),
                             _p(deltas), _p(weights_sum), _p(weights_edit_sum) if distill else None, _p(depth),
                             _p(depth_edit) if distill else None, _p(edit_occ) if distill else None, _p(image))
    out = (rays_alive, rays_t, weights_sum, depth, image)
    return out + (weights_edit_sum, depth_edit) if distill else out


def composite_rays_steps(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """composite_rays plus, per alive slot, the samples the ray completed in this round (orc_composite_rays_steps; test bookkeeping).
    Returns (rays_alive, rays_t, weights_sum, depth, image, steps_done)."""
    rays_alive, rays_t = _i32(rays_alive).copy(), _f32(rays_t).copy()
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    weights_sum, depth, image = _f32(weights_sum).copy(), _f32(depth).copy(), _f32(image).copy()
    steps = np.zeros(max(int(n_alive), 1), np.int32)
    lib().orc_composite_rays_steps(u32(n_alive), u32(n_step), f32c(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs), _p(deltas),
                                   _p(weights_sum), _p(depth), _p(image), _p(steps))
    return rays_alive, rays_t, weights_sum, depth, image, steps[:n_alive]


def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, align_corners=False):
    """Restates GridEncoder.__init__'s table sizing (gridencoder/grid.py:100-127). Returns (offsets int32[L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def grid_level_scales(L, S, H):
    out = np.empty(L, np.float32)
    lib().orc_grid_level_scales(u32(L), f32c(S), u32(H), _p(out))
    return out


def grid_encode_forward(inputs, embeddings, offsets, S, H, calc_grad_inputs=False, gridtype=0, align_corners=False, interp=0,
                        scales=None, out_layout=1):
    inputs, embeddings, offsets = _f32(inputs), _f32(embeddings), _i32(offsets)
    B, D = inputs.shape
    L, C_ = offsets.shape[0] - 1, embeddings.shape[1]
    out = np.empty((L, B, C_) if out_layout == 0 else (B, L * C_), np.float32)
    dy_dx = np.empty((B, L * D * C_), np.float32) if calc_grad_inputs else None
    scales = None if scales is None else _f32(scales)
    lib().orc_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(out), u32(B), u32(D), u32(C_), u32(L), f32c(S),
                                  u32(H), _p(dy_dx), u32(gridtype), C.c_int(int(align_corners)), u32(interp), _p(scales),
                                  C.c_int(out_layout))
    return (out, dy_dx) if calc_grad_inputs else out


def grid_encode_backward(grad, inputs, offsets, C_, S, H, dy_dx=None, gridtype=0, align_corners=False, interp=0, scales=None,
                         grad_layout=1):
    grad, inputs, offsets = _f32(grad), _f32(inputs), _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    ge = np.zeros((int(offsets[-1]), C_), np.float64)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    dy_dx = None if dy_dx is None else _f32(dy_dx)
    scales = None if scales is None else _f32(scales)
    lib().orc_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(ge), u32(B), u32(D), u32(C_), u32(L), f32c(S), u32(H),
                                   _p(dy_dx), _p(gi), u32(gridtype), C.c_int(int(align_corners)), u32(interp), _p(scales),
                                   C.c_int(grad_layout))
    return (ge, gi) if gi is not None else ge


def grad_total_variation(inputs, embeddings, grad, offsets, weight, S, H, gridtype=0, align_corners=False, scales=None):
    inputs, embeddings, offsets = _f32(inputs), _f32(embeddings), _i32(offsets)
    grad = _f32(grad).copy()
    B, D = inputs.shape
    L, C_ = offsets.shape[0] - 1, embeddings.shape[1]
    scales = None if scales is None else _f32(scales)
    lib().orc_grad_total_variation(_p(inputs), _p(embeddings), _p(grad), _p(offsets), f32c(weight), u32(B), u32(D), u32(C_),
                                   u32(L), f32c(S), u32(H), u32(gridtype), C.c_int(int(align_corners)), _p(scales))
    return grad


def ffmlp_forward(inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation=0, output_activation=6,
                  want_buffer=True):
    inputs, weights = _f32(inputs), _f32(weights)
    B = inputs.shape[0]
    out = np.empty((B, output_dim), np.float32)
    fb = np.empty((num_layers, B, hidden_dim), np.float32) if want_buffer else None
    lib().orc_ffmlp_forward(_p(inputs), _p(weights), u32(B), u32(input_dim), u32(output_dim), u32(hidden_dim), u32(num_layers),
                            u32(activation), u32(output_activation), _p(fb), _p(out))
    return (out, fb) if want_buffer else out


def ffmlp_backward(grad, inputs, weights, forward_buffer, input_dim, output_dim, hidden_dim, num_layers, activation=0,
                   calc_grad_inputs=False):
    grad, inputs, weights, forward_buffer = _f32(grad), _f32(inputs), _f32(weights), _f32(forward_buffer)
    B = inputs.shape[0]
    gw = np.zeros(weights.shape[0], np.float64)
    gi = np.empty((B, input_dim), np.float32) if calc_grad_inputs else None
    lib().orc_ffmlp_backward(_p(grad), _p(inputs), _p(weights), _p(forward_buffer), u32(B), u32(input_dim), u32(output_dim),
                             u32(hidden_dim), u32(num_layers), u32(activation), _p(gw), _p(gi))
    return gw, gi


def sh_encode(dirs, degree=4):
    dirs = _f32(dirs).reshape(-1, 3)
    out = np.empty((dirs.shape[0], degree * degree), np.float32)
    lib().orc_sh_encode(_p(dirs), u32(dirs.shape[0]), u32(degree), _p(out))
    return out
