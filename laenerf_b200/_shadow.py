"""fp16 shadows of fp32 parameters (row f-4 of SURVEY.md section 8) and their staleness rule.

`laenerf_b200.optim.AmpAdam` gives every parameter tensor a persistent fp16 copy (`owner._shadow_f16`) that the kernels read
under autocast instead of re-casting the fp32 tensor on every forward (the reference: gridencoder/grid.py:43-44,
ffmlp/ffmlp.py:23).  The optimizer kernels write both copies through raw pointers, which torch does not see; every OTHER
write to the fp32 parameter -- `load_state_dict`, `reset_parameters`, `ema.copy_to()`, a manual `p.data.copy_()` -- is an
in-place torch op and bumps `param._version`.  So the shadow is current iff the version recorded when it was last derived
still matches; otherwise it is re-derived (and, in ray-sharded training, the rank's fp32 master slice with it) before use --
the reference's "cast on every forward" semantics without the cast on every forward.
"""
from __future__ import annotations

import torch


def mark_current(owner, param) -> None:
    owner._shadow_version = param._version


def shadow_f16(owner, param):
    """The up-to-date fp16 shadow of `param`, or None when no optimizer keeps one."""
    sh = getattr(owner, "_shadow_f16", None)
    if sh is None:
        return None
    if param._version != getattr(owner, "_shadow_version", None):
        resync = getattr(owner, "_shadow_resync", None)
        with torch.no_grad():
            if resync is not None:
                resync()
            else:
                sh.copy_(param.data)
        owner._shadow_version = param._version
    return sh


def half_of(owner, param):
    """What the kernels gather from under fp16 autocast: the shadow when there is one, else a fresh cast."""
    sh = shadow_f16(owner, param)
    return sh if sh is not None else param.detach().half()
