"""Build libsanerf_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m sanerf_hq_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands in sanerf_hq_b200/lib/ (git-ignored,
but it ships to the GPU box with the gpurun snapshot).  No torch headers are involved: the
library's interface is the plain C ABI of include/sanerf_b200.h.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SANERF_LIB_VARIANT=<name> builds / loads sanerf_hq_b200/lib_<name>/ instead (kernel experiments: several variants of the
# library, compiled with different SANERF_NVCC_FLAGS, side by side in one gpurun snapshot)
_VARIANT = os.environ.get("SANERF_LIB_VARIANT", "")
LIBDIR = os.path.join(HERE, "lib_" + _VARIANT if _VARIANT else "lib")
SO = os.path.join(LIBDIR, "libsanerf_b200.so")
SOURCES = ["grid_encode.cu", "sh_encode.cu", "freq_encode.cu", "render.cu", "mlp_tc.cu", "heads.cu", "peer.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc.cuh"), os.path.join(CSRC, "grid_dev.cuh"), os.path.join(os.path.dirname(HERE), "include", "sanerf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "2"] + os.environ.get("SANERF_NVCC_FLAGS", "").split()


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = [os.path.join(LIBDIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    nvcc = _nvcc()

    def compile_one(so):
        s, o = so
        if force or _stale(o, [s] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            return True
        return False

    with ThreadPoolExecutor(max_workers=4) as ex:
        rebuilt = list(ex.map(compile_one, zip(srcs, objs)))
    if any(rebuilt) or _stale(SO, objs):
        cmd = [nvcc, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
