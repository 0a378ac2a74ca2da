#!/usr/bin/env python
"""Build the reference's OWN CUDA extensions for sm_100a into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Nothing under oracle/ is product code.  This recipe compiles the untouched sources where they lie
under /root/reference (nothing is copied into the repo) with one flag change: -std=c++14 -> -std=c++17
(torch 2.11 headers refuse C++14; see SURVEY.md header).  Outputs go ONLY to oracle/_ref/<name>/ which
is git-ignored but travels to the GPU box with the gpurun snapshot.

The built modules are used by
  * oracle/gen_golden.py      -- runs the real reference kernels on a B200 and freezes golden vectors
  * tests (-m gpu, optional)  -- side-by-side parity when oracle/_ref is present
  * bench.py (optional info)  -- times the reference extensions next to ours ("gpu_reference")
They are never imported by the laenerf_b200 package.

Usage:  python oracle/build_ref.py [raymarching gridencoder ffmlp shencoder]   (default: all four)
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


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent; nothing to do (the GPU box only uses prebuilt files)")
        sys.exit(0)
    want = sys.argv[1:] or ["raymarching", "gridencoder", "ffmlp", "shencoder"]
    for w in want:
        build("_" + w.lstrip("_"))
