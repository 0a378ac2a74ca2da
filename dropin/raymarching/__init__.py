"""`import raymarching` drop-in (reference: raymarching/__init__.py does `from .raymarching import *`)."""
from laenerf_b200.raymarching import *  # noqa: F401,F403
from laenerf_b200 import raymarching  # `from raymarching import raymarching` (editing/editgrid.py:3)
