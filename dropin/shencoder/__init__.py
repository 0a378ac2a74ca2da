"""`from shencoder import SHEncoder` drop-in (reference: shencoder/__init__.py)."""
from laenerf_b200.shencoder import SHEncoder, sh_encode  # noqa: F401
