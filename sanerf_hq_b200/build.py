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
# (source, object name, extra flags): render.cu is compiled twice -- the primary 16-warp flavour and the 20-warp flavour that only
# carries the SAM-frame kernels (see the head of render.cu)
SOURCES = [("grid_encode.cu", "grid_encode", []), ("sh_encode.cu", "sh_encode", []), ("freq_encode.cu", "freq_encode", []),
           ("render.cu", "render", []), ("render.cu", "render_w20", ["-DSANERF_RENDER_FLAVOUR=20"]), ("mlp_tc.cu", "mlp_tc", []),
           ("heads.cu", "heads", []), ("peer.cu", "peer", [])]
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
    units = [(os.path.join(CSRC, s), os.path.join(LIBDIR, o + ".o"), fl) for s, o, fl in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = [o for _, o, _ in units]
    nvcc = _nvcc()

    def compile_one(unit):
        s, o, fl = unit
        if force or _stale(o, [s] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + fl + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            return True
        return False

    with ThreadPoolExecutor(max_workers=4) as ex:
        rebuilt = list(ex.map(compile_one, units))
    if any(rebuilt) or _stale(SO, objs):
        cmd = [nvcc, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
