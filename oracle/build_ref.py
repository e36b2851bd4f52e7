"""Build recipe for oracle/_ref: the reference's OWN CUDA encoder kernels.

TEST INFRASTRUCTURE ONLY.  Compiles the reference sources *where they lie*
under /root/reference (nothing is copied into this repo) into pybind modules
    oracle/_ref/_gridencoder.so   <- gridencoder/src/{gridencoder.cu,bindings.cpp}
    oracle/_ref/_shencoder.so     <- shencoder/src/{shencoder.cu,bindings.cpp}
    oracle/_ref/_freqencoder.so   <- freqencoder/src/{freqencoder.cu,bindings.cpp}
with a flag-only change: -std=c++17 instead of the reference's -std=c++14
(gridencoder/backend.py:6-12), which torch 2.11 headers reject.  freqencoder
keeps its -use_fast_math (freqencoder/backend.py:9).

The modules need a GPU to *run*; they are used on the GPU box by
tests/test_ref_cuda_gpu.py / test_ref_cuda_bwd_gpu.py (new kernels vs the
reference's kernels), and -- together with the reference's byte-compiled Python
(oracle/stage_ref.py) through oracle/ref_runtime.py -- by the full-frame parity
tests, the reference-Trainer tests and bench.py's `ref_gpu_baseline`.  oracle/_ref/ is git-ignored but ships
with gpurun.  Takes ~15 min on 8 cores; run once:

    python oracle/build_ref.py            # all three
    python oracle/build_ref.py shencoder  # one
"""
import os
import shutil
import sys

REF = os.environ.get("SANERF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

MODS = {
    "gridencoder": ("_gridencoder", ["gridencoder.cu", "bindings.cpp"], []),
    "shencoder": ("_shencoder", ["shencoder.cu", "bindings.cpp"], []),
    "freqencoder": ("_freqencoder", ["freqencoder.cu", "bindings.cpp"], ["-use_fast_math"]),
}


def build(which):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    name, srcs, extra = MODS[which]
    bdir = os.path.join(OUT, "build_" + which)
    os.makedirs(bdir, exist_ok=True)
    load(
        name=name,
        sources=[os.path.join(REF, which, "src", s) for s in srcs],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__",
                           "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__"] + extra,
        build_directory=bdir,
        is_python_module=False,
        verbose=True,
    )
    shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
    shutil.rmtree(bdir, ignore_errors=True)
    print("built", os.path.join(OUT, name + ".so"))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("reference tree not present at", REF, "- nothing to build")
        sys.exit(0)
    for w in (sys.argv[1:] or list(MODS)):
        build(w)
