#!/usr/bin/env python
"""GPU diagnostic: FFMLP backward of liblaenerf_b200.so against a float64 numpy evaluation of the same math (with
the fp16 roundings the kernel applies), per weight block, for several batch sizes and shapes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from backends import OursBackend


def h16(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float64)


def truth(g, x, w, fb, in_dim, nl, calc_gi):
    blocks, off = [], 0
    blocks.append((off, 64, in_dim)); off += 64 * in_dim
    for _ in range(nl - 1):
        blocks.append((off, 64, 64)); off += 4096
    blocks.append((off, 16, 64))
    W = [w[o:o + r * c].reshape(r, c).astype(np.float64) for o, r, c in blocks]
    gw = np.zeros_like(w, dtype=np.float64)
    G = g.astype(np.float64)
    inputs = [x.astype(np.float64)] + [fb[l].astype(np.float64) for l in range(nl)]
    for m in range(nl, -1, -1):
        o, r, c = blocks[m]
        gw[o:o + r * c] = (G.T @ inputs[m]).reshape(-1)
        if m > 0:
            G = h16((G @ W[m]) * (inputs[m] > 0))
    gi = h16(G @ W[0]) if calc_gi else None
    return gw, gi, blocks


be = OursBackend()
rng = np.random.default_rng(0)
for in_dim, nl, B in ((32, 2, 128), (32, 2, 1024), (32, 3, 128), (32, 3, 256), (32, 3, 1024), (32, 3, 128 * 300), (64, 2, 512), (16, 4, 384)):
    std = np.sqrt(3 / 64)
    w = h16(rng.uniform(-std, std, 64 * (in_dim + 64 * (nl - 1) + 16))).astype(np.float32)
    x = h16(rng.standard_normal((B, in_dim)) * 0.5).astype(np.float32)
    g = h16(rng.standard_normal((B, 16)) * 1e-2).astype(np.float32)
    out, fb = be.ffmlp_fwd(x, w, in_dim, 16, 64, nl)
    gw, gi = be.ffmlp_bwd(g, x, w, fb, in_dim, 16, 64, nl, 0, True)
    tw, ti, blocks = truth(g, x, w, fb, in_dim, nl, True)
    msg = []
    for m, (o, r, c) in enumerate(blocks):
        a, b = gw[o:o + r * c].astype(np.float64), tw[o:o + r * c]
        msg.append(f"W{m}: {np.abs(a - b).max() / (np.abs(b).max() + 1e-30):.2e}")
    msg.append(f"gi: {np.abs(gi - ti).max() / (np.abs(ti).max() + 1e-30):.2e}")
    print(f"in={in_dim} nl={nl} B={B}: rel-to-max err " + "  ".join(msg), flush=True)
