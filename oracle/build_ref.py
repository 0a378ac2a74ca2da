#!/usr/bin/env python
"""Build the reference's OWN CUDA extensions for sm_100a into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Nothing under oracle/ is product code.  This recipe compiles the untouched sources where they lie
under /root/reference (nothing is copied into the repo) with one flag change: -std=c++14 -> -std=c++17
(torch 2.11 headers refuse C++14; see SURVEY.md header).  Outputs go ONLY to oracle/_ref/<name>/ which
is git-ignored but travels to the GPU box with the gpurun snapshot.

The built modules are used by
  * oracle/gen_golden.py      -- runs the real reference kernels on a B200 and freezes golden vectors
  * tests (-m gpu)            -- side-by-side parity when oracle/_ref is present (tests/backends.py RefBackend, tests/ref_stack.py)
  * bench.py                  -- times the reference stack on the same GPU next to ours ("gpu_reference")
They are never imported by the laenerf_b200 package.

`stage_python()` additionally stages the reference's own PYTHON callers of the hot path -- the four wrapper packages
(raymarching/, gridencoder/, ffmlp/, shencoder/), nerf/renderer.py, nerf/network_ff.py, nerf/network.py, encoding.py,
activation.py -- byte for byte into git-ignored oracle/_ref/py/ (SURVEY.md section 7.1), because /root/reference does not
exist on the GPU box.  Three tiny stub modules written HERE (not copied) stand in for imports the hot path never uses:
`trimesh` (renderer.py:2, a debug point-cloud viewer), `turtle` (ffmlp.py:2, an accidental import that needs tkinter) and
`nerf/utils.py` reduced to `custom_meshgrid` (the real one imports tensorboardX, lpips, mcubes, ... none of them installed).
tests/ref_stack.py puts that tree on sys.path either with the reference extensions ("reference" stack) or with dropin/
in front of it ("dropin" stack): the reference's callers then run unmodified on this library.

Usage:  python oracle/build_ref.py [raymarching gridencoder ffmlp shencoder py]   (default: all)
"""
import os
import sys
import time

REF = os.environ.get("LAENERF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# module name -> (reference sub-directory, sources, extra include dirs)
EXTS = {
    "_raymarching": ("raymarching", ["raymarching.cu", "bindings.cpp"], []),
    "_gridencoder": ("gridencoder", ["gridencoder.cu", "bindings.cpp"], []),
    "_shencoder": ("shencoder", ["shencoder.cu", "bindings.cpp"], []),
    "_ffmlp": ("ffmlp", ["ffmlp.cu", "bindings.cpp"],
               ["dependencies/cutlass/include", "dependencies/cutlass/tools/util/include"]),
}


def build(name: str) -> str:
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    sub, srcs, incs = EXTS[name]
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    nvcc_flags = [
        "-O3", "-std=c++17",
        "--expt-extended-lambda", "--expt-relaxed-constexpr",
        "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
        "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing",
    ]
    t0 = time.time()
    load(
        name=name,
        sources=[os.path.join(REF, sub, "src", s) for s in srcs],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=nvcc_flags,
        extra_include_paths=[os.path.join(REF, sub, i) for i in incs],
        build_directory=bdir,
        is_python_module=False,  # do not dlopen libcuda-dependent code on the CPU box; just build
        verbose=False,
    )
    so = os.path.join(bdir, name + ".so")
    print(f"[build_ref] {name}: {so} ({time.time() - t0:.0f}s)", flush=True)
    return so


PY_FILES = [
    "raymarching/__init__.py", "raymarching/raymarching.py", "raymarching/backend.py",
    "gridencoder/__init__.py", "gridencoder/grid.py", "gridencoder/backend.py",
    "ffmlp/__init__.py", "ffmlp/ffmlp.py", "ffmlp/backend.py",
    "shencoder/__init__.py", "shencoder/sphere_harmonics.py", "shencoder/backend.py",
    "nerf/renderer.py", "nerf/network_ff.py", "nerf/network.py", "encoding.py", "activation.py",
]

STUBS = {
    # written here, not copied: stand-ins for imports the hot path never executes
    "stubs/trimesh.py": "# stub: nerf/renderer.py:2 imports trimesh for plot_pointcloud (a debug viewer) only\n",
    "stubs/turtle.py": "# stub: ffmlp/ffmlp.py:2 does `from turtle import backward, forward` (unused; needs tkinter)\n"
                       "def backward(*a, **k):\n    raise NotImplementedError\n\n\ndef forward(*a, **k):\n    raise NotImplementedError\n",
    "pkgs/nerf/__init__.py": "",
    "pkgs/nerf/utils.py": "# stub of nerf/utils.py reduced to the one helper nerf/renderer.py imports (utils.py:43-48); the real module\n"
                          "# imports tensorboardX / lpips / mcubes / torch_ema, none of which is installed\n"
                          "import torch\n\n\ndef custom_meshgrid(*args):\n    return torch.meshgrid(*args, indexing='ij')\n",
}


def stage_python() -> str:
    """Copy the reference's hot-path Python callers untouched into oracle/_ref/py/pkgs and write the stubs."""
    import shutil
    dst_root = os.path.join(OUT, "py")
    for rel in PY_FILES:
        dst = os.path.join(dst_root, "pkgs", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for rel, text in STUBS.items():
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(text)
    print(f"[build_ref] staged {len(PY_FILES)} reference python files + {len(STUBS)} stubs under {dst_root}", flush=True)
    return dst_root


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent; nothing to do (the GPU box only uses prebuilt files)")
        sys.exit(0)
    want = sys.argv[1:] or ["raymarching", "gridencoder", "ffmlp", "shencoder", "py"]
    for w in want:
        if w == "py":
            stage_python()
        else:
            build("_" + w.lstrip("_"))
