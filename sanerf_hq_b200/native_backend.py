"""The reference's pybind `_backend` modules on top of libsanerf_b200 -- the B-native boundary of SURVEY.md 8b.

The reference's Python (gridencoder/grid.py:9-12, shencoder/sphere_harmonics.py:9-12, freqencoder/freq.py:9-12) does

    try:    import _gridencoder as _backend          # ahead-of-time build (setup.py)
    except ImportError: from .backend import _backend   # JIT build (backend.py)

and then calls 8 functions with tensors and scalars (gridencoder/src/bindings.cpp, shencoder/src/bindings.cpp,
freqencoder/src/bindings.cpp).  `install()` registers three modules with exactly those names and functions in `sys.modules`;
the reference's own `GridEncoder` / `SHEncoder` / `FreqEncoder` classes, its `NeRFNetwork` and its renderer then run
unmodified on the sm_100a kernels behind the C ABI (include/sanerf_b200.h):

    import sanerf_hq_b200.native_backend as nb; nb.install()     # before the first `import gridencoder`
    # ... the reference's main.py / its own modules, unchanged

Same argument order and meaning as the pybind functions; outputs are written in place; nothing is returned.  Differences a
maintainer should know: fp32 only (the reference force-disables fp16, main.py:217 -- half tensors raise instead of taking the
reference's half kernels), launches go to torch's CURRENT stream, and a failed launch raises RuntimeError (the reference never
checks).  There is no CPU path: a non-CUDA or non-contiguous tensor raises like the reference's CHECK_CUDA / CHECK_CONTIGUOUS.
"""
import sys
import types

import torch

from . import _lib


def _f32(*tensors, what):
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError(f"{what}: fp32 tensors only (got {t.dtype}); the B200 kernels have no half path")


def _go(what, tensors, call):
    _lib.require_cuda(*tensors, what=what)
    dev = next(t for t in tensors if t is not None).device
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(call(lib, _lib.stream_ptr()), what)
    _lib.count_launch()


# ---- _gridencoder (gridencoder/src/bindings.cpp:5-10; gridencoder.h:11-16) -----------------------------------------------
def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype, align_corners, interp):
    """inputs [B,D] in [0,1]; embeddings [sO,C]; offsets int32 [L+1]; outputs [L,B,C]; dy_dx None or [B,L*D*C]."""
    _f32(inputs, embeddings, outputs, dy_dx, what="grid_encode_forward")
    _go("grid_encode_forward", (inputs, embeddings, offsets, outputs, dy_dx),
        lambda lib, st: lib.sanerf_grid_encode_forward(_lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(outputs),
                                                       B, D, C, L, max_level, float(S), H, _lib.ptr(dy_dx), gridtype, int(align_corners),
                                                       interp, st))


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H, dy_dx, grad_inputs, gridtype,
                         align_corners, interp):
    """grad [L,B,C]; grad_embeddings [sO,C] accumulated into; dy_dx / grad_inputs None or [B,L*D*C] / [B,D]."""
    _f32(grad, inputs, embeddings, grad_embeddings, dy_dx, grad_inputs, what="grid_encode_backward")
    _go("grid_encode_backward", (grad, inputs, embeddings, offsets, grad_embeddings, dy_dx, grad_inputs),
        lambda lib, st: lib.sanerf_grid_encode_backward(_lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                                                        _lib.ptr(grad_embeddings), B, D, C, L, max_level, float(S), H, _lib.ptr(dy_dx),
                                                        _lib.ptr(grad_inputs), gridtype, int(align_corners), interp, st))


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    _f32(inputs, embeddings, grad, what="grad_total_variation")
    _go("grad_total_variation", (inputs, embeddings, grad, offsets),
        lambda lib, st: lib.sanerf_grad_total_variation(_lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(grad), _lib.ptr(offsets),
                                                        float(weight), B, D, C, L, float(S), H, gridtype, int(align_corners), st))


def grad_weight_decay(embeddings, grad, offsets, weight, B, C, L):
    _f32(embeddings, grad, what="grad_weight_decay")
    _go("grad_weight_decay", (embeddings, grad, offsets),
        lambda lib, st: lib.sanerf_grad_weight_decay(_lib.ptr(embeddings), _lib.ptr(grad), _lib.ptr(offsets), float(weight), B, C, L, st))


# ---- _shencoder (shencoder/src/bindings.cpp; shencoder.h:9-10) -----------------------------------------------------------
def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
    """inputs [B,D]; outputs [B,C^2]; C = degree; dy_dx None or [B, D*C^2]."""
    _f32(inputs, outputs, dy_dx, what="sh_encode_forward")
    _go("sh_encode_forward", (inputs, outputs, dy_dx),
        lambda lib, st: lib.sanerf_sh_encode_forward(_lib.ptr(inputs), _lib.ptr(outputs), B, D, C, _lib.ptr(dy_dx), st))


def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    _f32(grad, inputs, dy_dx, grad_inputs, what="sh_encode_backward")
    _go("sh_encode_backward", (grad, inputs, dy_dx, grad_inputs),
        lambda lib, st: lib.sanerf_sh_encode_backward(_lib.ptr(grad), _lib.ptr(inputs), B, D, C, _lib.ptr(dy_dx), _lib.ptr(grad_inputs), st))


# ---- _freqencoder (freqencoder/src/bindings.cpp; freqencoder.h:7-10) -----------------------------------------------------
def freq_encode_forward(inputs, B, D, deg, C, outputs):
    _f32(inputs, outputs, what="freq_encode_forward")
    _go("freq_encode_forward", (inputs, outputs),
        lambda lib, st: lib.sanerf_freq_encode_forward(_lib.ptr(inputs), B, D, deg, C, _lib.ptr(outputs), st))


def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
    _f32(grad, outputs, grad_inputs, what="freq_encode_backward")
    _go("freq_encode_backward", (grad, outputs, grad_inputs),
        lambda lib, st: lib.sanerf_freq_encode_backward(_lib.ptr(grad), _lib.ptr(outputs), B, D, deg, C, _lib.ptr(grad_inputs), st))


EXPORTS = {
    "_gridencoder": (grid_encode_forward, grid_encode_backward, grad_total_variation, grad_weight_decay),
    "_shencoder": (sh_encode_forward, sh_encode_backward),
    "_freqencoder": (freq_encode_forward, freq_encode_backward),
}


def modules():
    """{"_gridencoder": module, "_shencoder": module, "_freqencoder": module} with the reference's pybind function names."""
    out = {}
    for name, fns in EXPORTS.items():
        m = types.ModuleType(name)
        m.__doc__ = f"{name}: the reference's pybind module, served by libsanerf_b200 (sanerf_hq_b200.native_backend)"
        for fn in fns:
            setattr(m, fn.__name__, fn)
        out[name] = m
    return out


def install():
    """Register the three modules in sys.modules (the names the reference's `import _gridencoder as _backend` looks up first)."""
    _lib.load()            # fail now, loudly, if the library is missing
    mods = modules()
    sys.modules.update(mods)
    return mods
