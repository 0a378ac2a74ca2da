"""laenerf_b200 -- B200-native (sm_100a) implementation of LAENeRF's ray-marched NeRF step behind the reference's
own Python operator API.  Sub-modules mirror the reference packages one to one:

    laenerf_b200.raymarching   <- raymarching/raymarching.py
    laenerf_b200.gridencoder   <- gridencoder/grid.py
    laenerf_b200.ffmlp         <- ffmlp/ffmlp.py
    laenerf_b200.shencoder     <- shencoder/sphere_harmonics.py
    laenerf_b200.nerf          <- the callers (network_ff.NeRFNetwork, NeRFRenderer.run_cuda / run_cuda_distill /
                                  update_extra_state / mark_untrained_grid, the trainer's loss + optimizer recipe)
    laenerf_b200.style_encoder <- editing/style_encoder.py (LAENeRF recolouring / style network of the edit stage)
    laenerf_b200.optim         <- torch.optim.Adam + GradScaler as fused kernels, ray-sharded over NVLink peer memory
    laenerf_b200.parallel      <- one process per GPU: ray shards, image-tile shards, gathers

`dropin/` at the repository root holds top-level alias packages (`import raymarching`, `from gridencoder import
GridEncoder`, `from ffmlp import FFMLP`, `from shencoder import SHEncoder`) for use inside a LAENeRF checkout.
All compute goes through liblaenerf_b200.so (include/laenerf_b200.h); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"


def build(verbose: bool = False) -> str:
    from ._native import build as _b
    return _b(verbose)
