"""ctypes front-end of oracle/sanerf_oracle.c (CPU restatement of the reference's CUDA
encoder kernels).  TEST INFRASTRUCTURE ONLY -- see the header of sanerf_oracle.c.

Everything takes / returns CPU torch tensors (fp32, contiguous) so the oracle's torch
restatement of nerf/renderer.py (oracle/render_oracle.py) and the reference's own Python
(tests/golden/make_golden.py) can call it in place of the CUDA-only `_backend` modules.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsanerf_oracle.so")
_lib = None
_pool = None
_threads = os.cpu_count() or 1


def build(force=False):
    """Compile sanerf_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sanerf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def set_num_threads(n):
    global _threads, _pool
    _threads = max(1, int(n))
    _pool = None


def num_threads():
    return _threads


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u32, f32, vp, i32 = ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p, ctypes.c_int
        L.oracle_level_resolution.restype = u32
        L.oracle_level_resolution.argtypes = [u32, f32, u32]
        L.oracle_grid_encode_forward.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, u32, f32, u32, vp,
                                                 u32, i32, u32, u32, u32]
        L.oracle_grid_encode_backward.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, u32, f32, u32, u32, i32, u32]
        L.oracle_grid_input_backward.argtypes = [vp, vp, vp, u32, u32, u32, u32]
        L.oracle_grad_total_variation.argtypes = [vp, vp, vp, vp, f32, u32, u32, u32, u32, f32, u32, u32, i32]
        L.oracle_grad_weight_decay.argtypes = [vp, vp, vp, f32, u32, u32, u32]
        L.oracle_sh_encode_forward.argtypes = [vp, vp, u32, u32, u32, u32]
        L.oracle_freq_encode_forward.argtypes = [vp, u32, u32, u32, u32, vp]
        L.oracle_freq_encode_backward.argtypes = [vp, vp, u32, u32, u32, u32, vp]
        for f in ("oracle_grid_encode_forward", "oracle_grid_encode_backward", "oracle_grid_input_backward",
                  "oracle_grad_total_variation", "oracle_grad_weight_decay", "oracle_sh_encode_forward",
                  "oracle_freq_encode_forward", "oracle_freq_encode_backward"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t):
    assert t.device.type == "cpu"
    return t.detach().to(torch.float32).contiguous()


def _ranges(B):
    n = min(_threads, max(1, B // 2048))
    step = -(-B // n)
    return [(i, min(B, i + step)) for i in range(0, B, step)]


def _run(fn, B):
    global _pool
    rs = _ranges(B)
    if len(rs) <= 1:
        for r in rs:
            fn(*r)
        return
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=_threads)
    list(_pool.map(lambda r: fn(*r), rs))


def level_resolution(level, S, H):
    """Kernel-side fp32 resolution (gridencoder.cu:133) with glibc's exp2f."""
    return int(lib().oracle_level_resolution(int(level), float(np.float32(S)), int(H)))


def grid_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size):
    """Host-side row offsets table (gridencoder/grid.py:124-135), float64 like the reference."""
    offsets, off = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, res ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(off)
        off += n
    offsets.append(off)
    return torch.from_numpy(np.array(offsets, dtype=np.int32))


def grid_encode_forward(inputs, embeddings, offsets, B, D, C, L, max_level, S, H, dy_dx=None,
                        gridtype=0, align_corners=False, interp=0, outputs=None):
    """K1.  Same argument order/meaning as `_backend.grid_encode_forward` (gridencoder.h:12).
    Returns outputs [L,B,C] (allocated here when not passed in)."""
    inputs, embeddings = _f32(inputs), _f32(embeddings)
    offsets = offsets.to(torch.int32).contiguous()
    if outputs is None:
        outputs = torch.zeros(L, B, C, dtype=torch.float32)
    Lb = lib()
    S32 = float(np.float32(S))

    def go(b0, b1):
        Lb.oracle_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(outputs), B, D, C, L,
                                      max_level, S32, H, _p(dy_dx), gridtype, int(bool(align_corners)),
                                      interp, b0, b1)
    _run(go, B)
    return outputs


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H,
                         dy_dx=None, grad_inputs=None, gridtype=0, align_corners=False, interp=0):
    """K2 (+K3 when dy_dx is given).  grad [L,B,C]; grad_embeddings zero-filled by the caller."""
    grad, inputs = _f32(grad), _f32(inputs)
    offsets = offsets.to(torch.int32).contiguous()
    lib().oracle_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(grad_embeddings), B, D, C, L,
                                      max_level, float(np.float32(S)), H, gridtype, int(bool(align_corners)), interp)
    if dy_dx is not None:
        lib().oracle_grid_input_backward(_p(grad), _p(dy_dx), _p(grad_inputs), B, D, C, L)


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype=0,
                         align_corners=False):
    """K4, in place on grad."""
    inputs, embeddings = _f32(inputs), _f32(embeddings)
    offsets = offsets.to(torch.int32).contiguous()
    lib().oracle_grad_total_variation(_p(inputs), _p(embeddings), _p(grad), _p(offsets), float(weight), B, D, C,
                                      L, float(np.float32(S)), H, gridtype, int(bool(align_corners)))


def grad_weight_decay(embeddings, grad, offsets, weight, B, C, L):
    """K5, in place on grad."""
    embeddings = _f32(embeddings)
    offsets = offsets.to(torch.int32).contiguous()
    lib().oracle_grad_weight_decay(_p(embeddings), _p(grad), _p(offsets), float(weight), B, C, L)


def sh_encode_forward(inputs, B, degree, outputs=None):
    """K6 values.  inputs [B,3] already normalised."""
    inputs = _f32(inputs)
    if outputs is None:
        outputs = torch.empty(B, degree * degree, dtype=torch.float32)
    Lb = lib()
    _run(lambda b0, b1: Lb.oracle_sh_encode_forward(_p(inputs), _p(outputs), B, degree, b0, b1), B)
    return outputs


def freq_encode_forward(inputs, B, D, deg, C, outputs=None):
    inputs = _f32(inputs)
    if outputs is None:
        outputs = torch.empty(B, C, dtype=torch.float32)
    lib().oracle_freq_encode_forward(_p(inputs), B, D, deg, C, _p(outputs))
    return outputs


def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
    grad, outputs = _f32(grad), _f32(outputs)
    lib().oracle_freq_encode_backward(_p(grad), _p(outputs), B, D, deg, C, _p(grad_inputs))
    return grad_inputs


# ---- module-level helpers mirroring the Python encoder classes (forward only) ----------

def grid_encoder_apply(x, embeddings, offsets, per_level_scale, base_resolution, bound=1,
                       gridtype=0, align_corners=False, interp=0):
    """GridEncoder.forward (gridencoder/grid.py:151-168): x [...,D] in [-bound,bound] -> [..., L*C]."""
    x = (x + bound) / (2 * bound)
    prefix = list(x.shape[:-1])
    D = x.shape[-1]
    x = x.reshape(-1, D)
    B, L, C = x.shape[0], offsets.shape[0] - 1, embeddings.shape[1]
    S = np.log2(per_level_scale)
    out = grid_encode_forward(x, embeddings, offsets, B, D, C, L, L, S, base_resolution,
                              None, gridtype, align_corners, interp)
    return out.permute(1, 0, 2).reshape(prefix + [L * C])


def sh_encoder_apply(d, degree=4, size=1):
    """SHEncoder.forward (shencoder/sphere_harmonics.py:75-90)."""
    d = d / size
    d = d / torch.norm(d, dim=-1, keepdim=True)
    prefix = list(d.shape[:-1])
    d = d.reshape(-1, 3)
    out = sh_encode_forward(d, d.shape[0], degree)
    return out.reshape(prefix + [degree * degree])
