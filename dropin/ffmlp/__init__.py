"""`from ffmlp import FFMLP` drop-in (reference: ffmlp/__init__.py)."""
from laenerf_b200.ffmlp import FFMLP, ffmlp_forward, convert_activation  # noqa: F401
