#!/usr/bin/env python
"""GPU helper: evaluate every parity case with liblaenerf_b200.so and save the raw outputs to gpurun_out/ours/ for
offline analysis on the CPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from backends import OursBackend
from cases import CASES, run_case

out = os.path.join(ROOT, "gpurun_out", "ours")
os.makedirs(out, exist_ok=True)
be = OursBackend()
for name in (sys.argv[1:] or list(CASES)):
    try:
        np.savez_compressed(os.path.join(out, f"ours_{name}.npz"), **run_case(name, be))
        print("dumped", name)
    except Exception as e:
        print("FAILED", name, type(e).__name__, e)
