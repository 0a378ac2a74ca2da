"""`from gridencoder import GridEncoder` drop-in (reference: gridencoder/__init__.py)."""
from laenerf_b200.gridencoder import GridEncoder, grid_encode  # noqa: F401
