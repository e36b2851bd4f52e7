"""Build recipe for oracle/_ref/bytecode: the reference's OWN Python for the hot path, as bytecode.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, and reference SOURCES are never copied into
this repository.  What travels instead is a *build output*, exactly like oracle/_ref/*.so (the reference's .cu files
compiled where they lie by oracle/build_ref.py): this recipe byte-compiles the reference's unmodified .py files where
they lie under /root/reference into sourceless modules

    oracle/_ref/bytecode/nerf/{renderer,network,utils,trainer}.bin
    oracle/_ref/bytecode/{encoding,activation}.bin
    oracle/_ref/bytecode/{gridencoder,shencoder,freqencoder}/{__init__,<module>,backend}.bin

(git-ignored, shipped by gpurun; the extension is .bin because snapshot tools drop *.pyc -- the content is a regular CPython
3.12 bytecode file).  oracle/ref_runtime.py imports them (its own meta-path finder) next to oracle/_ref/_gridencoder.so etc., which
gives the GPU box the reference's own `NeRFNetwork.render` on the reference's own CUDA kernels: the GPU oracle of
SURVEY.md 8c/8d and the "reference's own CUDA-extension build" that BASELINE.json's north_star sets as the bar.

    python oracle/stage_ref.py
"""
import os
import py_compile
import sys

REF = os.environ.get("SANERF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "bytecode")
EXT = ".bin"

FILES = [
    "nerf/renderer.py", "nerf/network.py", "nerf/utils.py", "nerf/trainer.py",
    "encoding.py", "activation.py",
    "gridencoder/__init__.py", "gridencoder/grid.py", "gridencoder/backend.py",
    "shencoder/__init__.py", "shencoder/sphere_harmonics.py", "shencoder/backend.py",
    "freqencoder/__init__.py", "freqencoder/freq.py", "freqencoder/backend.py",
]


def stage(force=False):
    """Compile FILES -> OUT (sourceless layout: <name>.bin where <name>.py would be).  Returns the list of outputs."""
    outs = []
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + EXT)
        outs.append(dst)
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the name shown in tracebacks; UNCHECKED_HASH: no source is consulted at import time
        py_compile.compile(src, cfile=dst, dfile="reference/" + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return outs


def staged():
    return all(os.path.exists(os.path.join(OUT, rel[:-3] + EXT)) for rel in FILES)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("reference tree not present at", REF, "- nothing to stage")
        sys.exit(0)
    for o in stage(force="--force" in sys.argv):
        print("staged", os.path.relpath(o, os.path.dirname(OUT)))
