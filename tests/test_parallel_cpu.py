"""CPU tier: the N > 1 host logic with world_size-2 gloo process groups (ray-sharded gradient exchange, tile-sharded
render gather, shard arithmetic).  No kernels run; tensors are CPU tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from laenerf_b200.parallel import allreduce_gradients, gather_image, gather_tiles, init_distributed, shard_range, tile_shard_indices
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    # gradient exchange: one "hash-grid" sized tensor (>= 2^20 elements -> own all-reduce) and two small MLP tensors
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(1 << 20, 2)), torch.nn.Parameter(torch.zeros(7168)), torch.nn.Parameter(torch.zeros(11264))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params.append(torch.nn.Parameter(torch.zeros(3)))  # no grad: must be skipped
    allreduce_gradients(params, w, average=True)
    want = [sum(range(1, world + 1)) / world * (i + 1) for i in range(3)]
    ok_grad = all(torch.allclose(p.grad, torch.full_like(p, want[i])) for i, p in enumerate(params[:3]))
    # tile-sharded render: contiguous ray ranges, ragged split, final gather reproduces the full image on every rank
    n = 1001
    full = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    lo, hi = shard_range(n, r, w)
    img = gather_image(full[lo:hi].clone(), n, r, w)
    # round-robin image tiles (ragged 50 x 70 image, 16-pixel tiles): the gather restores pixel order on every rank
    Hh, Ww = 50, 70
    pix = torch.arange(Hh * Ww * 3, dtype=torch.float32).view(Hh * Ww, 3)
    mine = tile_shard_indices(Hh, Ww, r, w, tile=16)
    img2 = gather_tiles(pix[mine].clone(), Hh, Ww, r, w, tile=16)
    out[rank] = (ok_grad, bool(torch.equal(img, full)) and bool(torch.equal(img2, pix)), (lo, hi))
    dist.destroy_process_group()


def test_world2_gloo_gradient_exchange_and_render_gather():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    assert all(v[0] and v[1] for v in out.values()), dict(out)
    assert out[0][2] == (0, 501) and out[1][2] == (501, 1001)


def test_shard_range_covers_everything_once():
    from laenerf_b200.parallel import shard_range
    for n in (0, 1, 7, 640000, 190512):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_tile_shards_partition_the_image_and_balance():
    from laenerf_b200.parallel import tile_shard_indices
    for H, W, tile in ((800, 800, 32), (378, 504, 32), (519, 779, 16), (5, 7, 4)):
        for w in (1, 2, 4, 8):
            parts = [tile_shard_indices(H, W, r, w, tile) for r in range(w)]
            allidx = torch.cat(parts)
            assert allidx.numel() == H * W and torch.equal(torch.sort(allidx).values, torch.arange(H * W))
            if H * W > 10000:
                sizes = [p.numel() for p in parts]
                assert max(sizes) - min(sizes) <= 2 * tile * tile * ((W + tile - 1) // tile)
