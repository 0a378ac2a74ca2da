"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` over NCCL/NVLink.

The reference is single-GPU (its DistributedDataParallel hooks are dead code, nerf/utils.py:380-382); SURVEY.md
section 8(e) defines what the hot path needs:
  * training = data parallel over RAYS: every rank marches/encodes/composites its own ray batch against replicated
    parameters and a replicated occupancy bitfield; the one exchange step is the sum of the hash-grid gradient and
    of the two flat MLP weight gradients, followed by the identical Adam step on every rank;
  * rendering = image tiles per rank, no collective except the final gather.  Tiles are small (32 x 32 pixels) and dealt
    round-robin: contiguous row ranges put the object in the middle ranks and empty sky in the others (measured on the
    lego-shape frame: the first 80 000 of 640 000 rays need 1 round and 0.8 ms, the whole frame 91 rounds and 19 ms).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None):
    """Returns (rank, world_size, local_rank); initialises the default process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) share of n units (rays of an image, views of a dataset) for `rank`; sizes differ by <= 1."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_shard_indices(H: int, W: int, rank: int, world: int, tile: int = 32) -> torch.Tensor:
    """Flat pixel (= ray) indices of the image tiles `rank` renders: tile t, counted row-major over the
    ceil(H/tile) x ceil(W/tile) tile grid, belongs to rank t % world.  Pixels stay tile-major (row-major inside a tile), so
    neighbouring lanes still march neighbouring rays.  Every rank can compute every other rank's list: the final gather needs
    no index exchange."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    t = torch.arange(ty * tx)
    mine = t[t % world == rank]
    y0, x0 = (mine // tx) * tile, (mine % tx) * tile
    dy, dx = torch.meshgrid(torch.arange(tile), torch.arange(tile), indexing="ij")
    yy = y0[:, None, None] + dy[None]
    xx = x0[:, None, None] + dx[None]
    ok = (yy < H) & (xx < W)
    return (yy * W + xx)[ok].reshape(-1)


_tile_plans = {}  # (H, W, world, tile, device) -> (maxlen, inverse permutation on the device)


def _tile_plan(H: int, W: int, world: int, tile: int, device):
    """Where every pixel of the full image sits in the rank-major buffer an all_gather of equal-sized padded pieces produces:
    computed once per image geometry and kept on the device (the gather itself is then one collective + one index_select)."""
    key = (H, W, world, tile, str(device))
    plan = _tile_plans.get(key)
    if plan is None:
        idx = [tile_shard_indices(H, W, r, world, tile) for r in range(world)]
        maxlen = max(i.numel() for i in idx)
        inv = torch.empty(H * W, dtype=torch.long)
        for r, i in enumerate(idx):
            inv[i] = r * maxlen + torch.arange(i.numel())
        plan = (maxlen, inv.to(device))
        _tile_plans[key] = plan
    return plan


def gather_tiles(local: torch.Tensor, H: int, W: int, rank: int, world: int, tile: int = 32, group=None):
    """Final gather of a tile-sharded render: `local` holds this rank's pixels in tile_shard_indices order; returns the full
    [H*W, ...] image on every rank: one all_gather of equal-sized padded pieces into a flat buffer, then one index_select with
    the cached inverse permutation."""
    if world <= 1:
        return local
    maxlen, inv = _tile_plan(H, W, world, tile, local.device)
    tail = tuple(local.shape[1:])
    if local.shape[0] == maxlen:
        piece = local.contiguous()
    else:
        piece = torch.zeros((maxlen,) + tail, dtype=local.dtype, device=local.device)
        piece[: local.shape[0]] = local
    flat = torch.empty((world * maxlen,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, piece, group=group)
    return flat.index_select(0, inv)


def allreduce_gradients(params, world: int, group=None, average: bool = True):
    """Sum (or mean) the gradients of `params` over ranks, in place, largest tensor first so that the 49 MB hash-grid
    gradient is in flight while the small MLP gradients are coalesced into one extra call."""
    if world <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    grads.sort(key=lambda g: -g.numel())
    big = [g for g in grads if g.numel() >= (1 << 20)]
    small = [g for g in grads if g.numel() < (1 << 20)]
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in big]
    if small:
        flat = torch.cat([g.reshape(-1).float() for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    for w in works:
        w.wait()
    if average:
        for g in grads:
            g.div_(world)


def allreduce_tensors(tensors, world: int, group=None, average: bool = True):
    """The same exchange step on explicit gradient buffers (laenerf_b200.optim.AmpAdam keeps fp16 ones: the hash-grid
    gradient crosses NVLink as 24.5 MB instead of 49 MB).  Mean over ranks: NCCL averages inside the collective
    (ReduceOp.AVG, no fp16 overflow from summing `world` loss-scaled gradients); other backends sum, then divide."""
    if world <= 1 or not tensors:
        return
    nccl = dist.get_backend(group) == "nccl"
    op = dist.ReduceOp.AVG if (average and nccl) else dist.ReduceOp.SUM
    tensors = sorted(tensors, key=lambda g: -g.numel())
    works = [dist.all_reduce(g, op=op, group=group, async_op=True) for g in tensors]
    for w in works:
        w.wait()
    if average and not nccl:
        for g in tensors:
            g.div_(world)


def broadcast_occupancy(model, src: int = 0, group=None):
    """Keep the occupancy state identical on every rank (the reference's update uses RNG, renderer.py:590-620)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model.density_bitfield, src=src, group=group)
        dist.broadcast(model.density_grid, src=src, group=group)


def gather_image(local: torch.Tensor, n_total: int, rank: int, world: int, group=None):
    """Final gather of a range-sharded render (contiguous ray ranges): `local` is this rank's [hi-lo, ...] slice; returns the full
    [n_total, ...] tensor on every rank (what nerf/utils.py:1560-1566 does with all_gather)."""
    if world <= 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
