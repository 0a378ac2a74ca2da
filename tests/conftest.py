import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    import numpy as np
    path = os.path.join(GOLDEN, f"ref_{name}.npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def golden_scales(gold):
    """{(L, H): scales} from the 'in_scales_L_H' entries of a golden file."""
    out = {}
    for k, v in (gold or {}).items():
        if k.startswith("in_scales_"):
            _, _, L, H = k.split("_")
            out[(int(L), int(H))] = v
    return out


@pytest.fixture(scope="session")
def oracle_backend():
    from backends import OracleBackend
    return OracleBackend()


@pytest.fixture(scope="session")
def ours_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from backends import OursBackend
    return OursBackend()
