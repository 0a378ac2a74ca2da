"""Synthetic procedural scenes of the shapes BASELINE.json names (no dataset / network access).

A scene here is what the hot path consumes: an occupancy bitfield in the reference's layout
(`density_bitfield`, uint8 [C*H^3/8], Morton order inside each cascade -- nerf/renderer.py:556-649,
raymarching/src/raymarching.cu:267-289) plus pinhole cameras producing rays exactly like
`get_rays` (nerf/utils.py:62-153).  Everything is numpy on the host and seeded; see SURVEY.md 8(d).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

GRID_H = 128  # the reference hard-codes grid_size = 128 (nerf/renderer.py:92)


def _expand_bits(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d_np(x, y, z) -> np.ndarray:
    """10-bit/axis Morton code, the layout of density_grid (raymarching.cu:56-71)."""
    with np.errstate(over="ignore"):
        return _expand_bits(np.asarray(x)) | (_expand_bits(np.asarray(y)) << np.uint32(1)) | (_expand_bits(np.asarray(z)) << np.uint32(2))


def packbits_np(grid: np.ndarray, thresh: float) -> np.ndarray:
    """bit i of byte n = grid.flat[8n+i] > thresh (raymarching.cu:281-288)."""
    bits = (grid.reshape(-1, 8) > np.float32(thresh)).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


@dataclass
class Scene:
    name: str
    bound: float
    cascade: int            # C = 1 + ceil(log2(bound))
    H: int
    W: int
    focal: float
    min_near: float
    density_grid: np.ndarray     # float32 [C, 128^3], Morton order
    density_bitfield: np.ndarray  # uint8 [C*128^3/8]
    poses: np.ndarray            # float32 [P, 4, 4] cam2world (camera looks along +z of its frame)
    density_thresh: float = 10.0
    dt_gamma: float = 0.0
    max_steps: int = 1024

    @property
    def aabb(self) -> np.ndarray:
        b = self.bound
        return np.array([-b, -b, -b, b, b, b], dtype=np.float32)

    @property
    def intrinsics(self):
        return (self.focal, self.focal, self.W / 2.0, self.H / 2.0)

    def occupancy_fraction(self) -> float:
        return float(np.unpackbits(self.density_bitfield).mean())


def _boxes_inside(p: np.ndarray, boxes) -> np.ndarray:
    inside = np.zeros(p.shape[:-1], dtype=bool)
    for lo, hi in boxes:
        inside |= np.all((p >= np.asarray(lo, np.float32)) & (p <= np.asarray(hi, np.float32)), axis=-1)
    return inside


def _grid_from_solid(inside_fn, C: int, bound: float, rng: np.random.Generator, H: int = GRID_H) -> np.ndarray:
    """density_grid[c, morton(x,y,z)] = 20*inside(cell centre) + U(0,1); cascade c spans [-min(2^c,bound), +...]."""
    grid = np.zeros((C, H ** 3), dtype=np.float32)
    idx = np.arange(H, dtype=np.uint32)
    X, Y, Z = np.meshgrid(idx, idx, idx, indexing="ij")
    mort = morton3d_np(X.ravel(), Y.ravel(), Z.ravel()).astype(np.int64)
    centres = (np.stack([X.ravel(), Y.ravel(), Z.ravel()], -1).astype(np.float32) + 0.5) / H * 2.0 - 1.0
    for c in range(C):
        mip_bound = min(2.0 ** c, bound)
        inside = inside_fn(centres * np.float32(mip_bound), half_cell=mip_bound / H)
        vals = 20.0 * inside.astype(np.float32) + rng.random(H ** 3, dtype=np.float32)
        grid[c, mort] = vals
    return grid


def _look_at(eye: np.ndarray, target: np.ndarray) -> np.ndarray:
    fwd = target - eye
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0]) if abs(fwd[2]) < 0.99 else np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, eye
    return pose


def make_scene(name: str = "lego", seed: int = 0, n_poses: int = 100) -> Scene:
    rng = np.random.default_rng(seed)
    if name == "lego":
        # blender lego shape: 800x800, camera_angle_x 0.6911, radius 4.03*scale(0.8), bound 1 (configs_nerf_synthetic/lego.sh)
        bound, W, H = 1.0, 800, 800
        focal = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
        boxes = [((-0.55, -0.30, -0.35), (0.55, 0.30, -0.25)),   # base plate
                 ((-0.50, -0.25, -0.25), (0.10, 0.25, -0.05)),   # chassis
                 ((-0.45, -0.20, -0.05), (-0.05, 0.20, 0.20)),   # cabin
                 ((0.10, -0.08, -0.25), (0.55, 0.08, -0.12)),    # arm base
                 ((0.30, -0.06, -0.12), (0.42, 0.06, 0.30)),     # arm upright
                 ((0.36, -0.20, 0.22), (0.60, 0.20, 0.34)),      # bucket
                 ((-0.55, -0.34, -0.40), (-0.25, -0.26, -0.20)), ((-0.55, 0.26, -0.40), (-0.25, 0.34, -0.20)),  # tracks
                 ((0.05, -0.34, -0.40), (0.45, -0.26, -0.20)), ((0.05, 0.26, -0.40), (0.45, 0.34, -0.20)),
                 ((-0.30, -0.12, 0.20), (-0.15, 0.12, 0.32)),    # roof light
                 ((-0.10, -0.28, -0.05), (0.05, -0.22, 0.10))]   # side box

        solid_scale = np.float32(1.7)  # tuned so a 4096-ray batch yields ~2^18 samples (~60 per ray)

        def inside_fn(p, half_cell):
            p = p / solid_scale
            slab = (np.abs(p[:, 2] - 0.5 * p[:, 0] - 0.05) < 0.03) & (np.abs(p[:, 0]) < 0.5) & (np.abs(p[:, 1]) < 0.1)
            return _boxes_inside(p, boxes) | slab

        C, min_near = 1, 0.2
        radius = 4.03 * 0.8
        poses = []
        for _ in range(n_poses):
            th = rng.uniform(0, 2 * math.pi)
            ph = rng.uniform(0.05, 0.5 * math.pi - 0.05)
            eye = radius * np.array([math.cos(th) * math.cos(ph), math.sin(th) * math.cos(ph), math.sin(ph)])
            poses.append(_look_at(eye, np.zeros(3)))
    elif name == "flower":
        # llff flower shape: 504x378, bound 2, scale 0.02 + offset (0,0,1.5) => cameras bunched near (0,0,1.5)
        bound, W, H = 2.0, 504, 378
        focal = 0.82 * W
        cl = np.random.default_rng(seed + 1)
        boxes = []
        for _ in range(14):
            c = np.array([cl.uniform(-0.9, 0.9), cl.uniform(-0.7, 0.7), cl.uniform(-1.0, 0.6)])
            h = cl.uniform(0.08, 0.3, size=3)
            boxes.append((c - h, c + h))
        boxes.append(((-1.8, -1.8, -1.9), (1.8, 1.8, -1.7)))  # background wall behind the cluster

        def inside_fn(p, half_cell):
            return _boxes_inside(p, boxes)

        C, min_near = 2, 0.2
        poses = []
        for _ in range(max(n_poses, 1)):
            eye = np.array([0.0, 0.0, 1.5]) + rng.uniform(-0.08, 0.08, size=3)
            poses.append(_look_at(eye, np.array([0.0, 0.0, -0.5]) + rng.uniform(-0.05, 0.05, size=3)))
    elif name == "bonsai":
        # mip-NeRF-360 bonsai shape, BASELINE bound 16 => 5 cascades; central object + sparse far shell
        bound, W, H = 16.0, 779, 519
        focal = 0.9 * W
        cl = np.random.default_rng(seed + 2)
        boxes = [((-0.35, -0.35, -0.4), (0.35, 0.35, -0.2)), ((-0.08, -0.08, -0.2), (0.08, 0.08, 0.25)),
                 ((-0.3, -0.3, 0.2), (0.3, 0.3, 0.5))]
        far = []
        for _ in range(40):
            d = cl.normal(size=3)
            d /= np.linalg.norm(d)
            c = d * cl.uniform(3.0, 14.0)
            h = cl.uniform(0.3, 1.2, size=3)
            far.append((c - h, c + h))

        def inside_fn(p, half_cell):
            return _boxes_inside(p, boxes) | _boxes_inside(p, far)

        C, min_near = 5, 0.05
        poses = []
        for _ in range(n_poses):
            th = rng.uniform(0, 2 * math.pi)
            eye = np.array([0.64 * math.cos(th), 0.64 * math.sin(th), rng.uniform(0.1, 0.4)])
            poses.append(_look_at(eye, np.zeros(3)))
    else:
        raise ValueError(f"unknown scene {name!r}")
    grid = _grid_from_solid(inside_fn, C, bound, rng)
    bitfield = packbits_np(grid, 10.0)
    return Scene(name=name, bound=bound, cascade=C, H=H, W=W, focal=float(focal), min_near=min_near,
                 density_grid=grid, density_bitfield=bitfield, poses=np.stack(poses).astype(np.float32))


def get_rays_np(pose: np.ndarray, intrinsics, H: int, W: int, N: int = -1, rng: np.random.Generator | None = None,
                inds: np.ndarray | None = None):
    """Pinhole rays of one camera as nerf/utils.py:62-153 builds them (pixel centres +0.5, z = 1, normalised,
    rotated by pose[:3,:3]); N > 0 draws N random pixels with replacement (utils.py:107-109)."""
    fx, fy, cx, cy = intrinsics
    if inds is None:
        if N > 0:
            inds = (rng or np.random.default_rng(0)).integers(0, H * W, size=min(N, H * W))
        else:
            inds = np.arange(H * W)
    i = (inds % W).astype(np.float32) + 0.5
    j = (inds // W).astype(np.float32) + 0.5
    d = np.stack([(i - np.float32(cx)) / np.float32(fx), (j - np.float32(cy)) / np.float32(fy), np.ones_like(i)], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = (d @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, rays_d, inds
