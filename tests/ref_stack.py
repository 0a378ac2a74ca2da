"""TEST INFRASTRUCTURE: the reference's OWN Python callers of the hot path (nerf/renderer.py `NeRFRenderer.run_cuda`,
`run_cuda_distill`, `update_extra_state`; nerf/network_ff.py and nerf/network.py `NeRFNetwork`; encoding.py; activation.py;
the four wrapper packages), staged untouched under oracle/_ref/py/ by oracle/build_ref.py, importable in two flavours:

    "reference"  the wrapper packages of the reference on the reference's own extensions (oracle/_ref/_*/_*.so)
    "dropin"     the SAME nerf/ + encoding.py + activation.py files, with dropin/ in front of the path: `import raymarching`,
                 `from gridencoder import GridEncoder`, `from ffmlp import FFMLP`, `from shencoder import SHEncoder` resolve
                 to laenerf_b200 -- the drop-in claim of INTEGRATION.md, executed

Both flavours use the same top-level module names, so each lives in its own snapshot of `sys.modules`: `use(kind)` swaps the
snapshot in, `load(kind)` returns a namespace of the flavour's modules.  Models must be CONSTRUCTED inside `use(kind)`
(encoding.get_encoder imports gridencoder / shencoder at call time); calling them afterwards needs no context.

Nothing here is imported by the laenerf_b200 package.  bench.py uses it for the `gpu_reference` record only.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "oracle", "_ref")
_PY = os.path.join(_REF, "py")
_EXTS = ("_raymarching", "_gridencoder", "_ffmlp", "_shencoder")
# top-level names that differ between the flavours (the compiled _ext modules are shared: only "reference" imports them)
_SWAPPED = ("raymarching", "gridencoder", "ffmlp", "shencoder", "encoding", "activation", "nerf", "trimesh", "turtle")
_snapshots: dict[str, dict] = {"reference": {}, "dropin": {}}


def available(kind: str = "reference") -> bool:
    if not os.path.exists(os.path.join(_PY, "pkgs", "nerf", "renderer.py")):
        return False
    if kind == "reference":
        return all(os.path.exists(os.path.join(_REF, m, m + ".so")) for m in _EXTS)
    return True


def _paths(kind: str):
    pk, st = os.path.join(_PY, "pkgs"), os.path.join(_PY, "stubs")
    if kind == "reference":
        return [os.path.join(_REF, m) for m in _EXTS] + [pk, st]
    if kind == "dropin":
        return [os.path.join(ROOT, "dropin"), pk, st]
    raise ValueError(kind)


def _mine(name: str) -> bool:
    return name.split(".", 1)[0] in _SWAPPED


@contextlib.contextmanager
def use(kind: str):
    """Make `kind`'s modules the ones `import` sees for the duration of the block."""
    if not available(kind):
        raise RuntimeError(f"reference stack '{kind}' is not staged: run `python oracle/build_ref.py` where /root/reference exists")
    outer = {k: sys.modules.pop(k) for k in list(sys.modules) if _mine(k)}
    sys.modules.update(_snapshots[kind])
    saved_path = list(sys.path)
    sys.path[:0] = _paths(kind)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)  # torch.cuda.amp.custom_fwd deprecation in the reference files
            yield
    finally:
        _snapshots[kind] = {k: sys.modules.pop(k) for k in list(sys.modules) if _mine(k)}
        sys.modules.update(outer)
        sys.path[:] = saved_path


def load(kind: str) -> types.SimpleNamespace:
    """Namespace with the flavour's modules: raymarching, gridencoder, ffmlp, shencoder, encoding, activation, renderer,
    network_ff, network."""
    with use(kind):
        ns = types.SimpleNamespace(kind=kind)
        for short, name in (("raymarching", "raymarching"), ("gridencoder", "gridencoder"), ("ffmlp", "ffmlp"),
                            ("shencoder", "shencoder"), ("encoding", "encoding"), ("activation", "activation"),
                            ("renderer", "nerf.renderer"), ("network_ff", "nerf.network_ff"), ("network", "nerf.network")):
            setattr(ns, short, importlib.import_module(name))
    return ns


def make_model(kind: str, variant: str = "ff", device=None, quiet: bool = True, **kw):
    """nerf/network_ff.py (`--ff`, variant "ff") or nerf/network.py (the default `-O` stack, variant "default") NeRFNetwork of
    the reference, constructed as main_nerf.py:130-143 does with cuda_ray=True."""
    ns = load(kind)
    mod = ns.network_ff if variant == "ff" else ns.network
    kw.setdefault("cuda_ray", True)
    with use(kind):
        if quiet:
            with contextlib.redirect_stdout(open(os.devnull, "w")):
                model = mod.NeRFNetwork(**kw)
        else:
            model = mod.NeRFNetwork(**kw)
    if device is not None:
        model = model.to(device)
    return model


def copy_state(dst, src):
    """Copy encoder table, MLP weights and occupancy state between any two of {reference-stack model, dropin-stack model,
    laenerf_b200.nerf.NeRFNetwork}: they share parameter / buffer names by construction (the drop-in contract)."""
    import torch
    sd = src.state_dict()
    own = dst.state_dict()
    with torch.no_grad():
        for k, v in own.items():
            if k in sd and sd[k].shape == v.shape:
                v.copy_(sd[k])
    for attr in ("mean_density", "iter_density", "mean_count", "local_step"):
        if hasattr(src, attr) and hasattr(dst, attr):
            try:
                setattr(dst, attr, getattr(src, attr))
            except AttributeError:
                pass
    return dst
