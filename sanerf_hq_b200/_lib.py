"""ctypes binding of libsanerf_b200.so -- the C ABI declared in include/sanerf_b200.h.

The product path has NO fallback: if the shared object is missing (and cannot be built because
nvcc is absent) importing an operator raises, and calling an operator on a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

from . import build as _build

_u32, _f32, _vp, _i32 = ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p, ctypes.c_int
MAX_LEVELS = 16
MAX_PEERS = 8
PEER_HANDLE_BYTES = 64


class GridT(ctypes.Structure):
    _fields_ = [("embeddings", _vp), ("num_levels", _u32), ("level_dim", _u32),
                ("offset", _u32 * (MAX_LEVELS + 1)), ("res", _u32 * MAX_LEVELS)]


class ModelT(ctypes.Structure):
    _fields_ = [("prop_grid", GridT * 2), ("prop_w0", _vp * 2), ("prop_w1", _vp * 2),
                ("grid", GridT), ("grid_w", _vp * 3), ("grid_hidden", _u32),
                ("view_w", _vp * 3), ("view_hidden", _u32),
                ("s_grid", GridT), ("sam_w", _vp * 5), ("sam_b", _vp * 5), ("sam_ln_w", _vp), ("sam_ln_b", _vp),
                ("m_grid", GridT), ("mask_w", _vp * 3), ("n_inst", _u32),
                ("aabb", _f32 * 6), ("min_near", _f32), ("grid_bound", _f32), ("contract", _u32),
                ("last_sample_opaque", _u32), ("u65", _vp), ("u33", _vp)]


class RenderArgsT(ctypes.Structure):
    _fields_ = [("rays_o", _vp), ("rays_d", _vp), ("N", _u32), ("cam_near_far", _vp), ("cam_near_far_rows", _u32),
                ("bg_color", _vp), ("bg_rows", _u32), ("bg_scalar", _f32),
                ("image", _vp), ("depth", _vp), ("weights_sum", _vp), ("sam_in", _vp), ("mask_in", _vp),
                ("mask_in_tiled", _u32), ("inds0", _vp), ("inds1", _vp), ("weights2", _vp), ("sigma2", _vp), ("bins2", _vp), ("f_image", _vp),
                ("cam_w", _u32), ("cam_ray0", _u32), ("cam_intrinsics", _f32 * 4), ("cam_pose", _f32 * 12), ("tile_w", _u32), ("image_u8", _vp),
                ("n_peer_out", _u32), ("peer_image", _vp * MAX_PEERS), ("peer_depth", _vp * MAX_PEERS), ("peer_weights_sum", _vp * MAX_PEERS),
                ("max_ctas", _u32), ("noise0", _vp), ("noise1", _vp), ("noise2", _vp), ("workspace", _vp)]


# name -> argtypes (restype is int for all but the two noted)
PROTOTYPES = {
    "sanerf_grid_encode_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _u32, _i32, _u32, _vp],
    "sanerf_grid_encode_forward_fused": [_vp, _f32, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _vp],
    "sanerf_grid_encode_backward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _vp, _u32, _i32, _u32, _vp],
    "sanerf_grid_encode_backward_fused": [_vp, _vp, _f32, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _u32, _vp],
    "sanerf_grad_total_variation": [_vp, _vp, _vp, _vp, _f32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _vp],
    "sanerf_grad_weight_decay": [_vp, _vp, _vp, _f32, _u32, _u32, _u32, _vp],
    "sanerf_grid_level_resolutions": [_vp, _u32, _f32, _u32, _vp],
    "sanerf_sh_encode_forward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "sanerf_sh_encode_backward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
    "sanerf_freq_encode_forward": [_vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "sanerf_freq_encode_backward": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "sanerf_render": [ctypes.POINTER(ModelT), ctypes.POINTER(RenderArgsT), _vp],
    "sanerf_render_workspace_bytes": [],
    "sanerf_sample_pdf": [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
    "sanerf_mlp3_tc": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp],
    "sanerf_mask_head_workspace_bytes": [],
    "sanerf_mask_head": [_vp, _vp, ctypes.POINTER(GridT), _vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp],
    "sanerf_samvit_mlp_workspace_bytes": [],
    "sanerf_samvit_mlp": [_vp, _vp * 5, _vp * 5, _vp, _vp, _u32, _vp, _vp, _vp],
    "sanerf_samvit_mlp_layout": [_vp, _vp * 5, _vp * 5, _vp, _vp, _u32, _vp, _vp, _u32, _vp],
    "sanerf_feature_resize_nchw": [_vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp],
    "sanerf_peer_alloc": [ctypes.c_size_t, ctypes.POINTER(_vp)],
    "sanerf_peer_free": [_vp],
    "sanerf_peer_export": [_vp, ctypes.c_char_p],
    "sanerf_peer_open": [ctypes.c_char_p, ctypes.POINTER(_vp)],
    "sanerf_peer_close": [_vp],
    "sanerf_peer_push": [ctypes.POINTER(_vp), _vp, ctypes.c_size_t, _u32, ctypes.POINTER(_vp)],
    "sanerf_peer_barrier": [ctypes.POINTER(_vp), _u32, _u32, _u32, _f32, _vp, _vp],
    "sanerf_abi_version": [],
    "sanerf_error_string": [_i32],
}
_lib = None


def lib_path():
    return _build.SO


def load():
    """dlopen the in-tree library (building it first if sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.SO
    try:
        path = _build.build()
    except Exception as e:  # no nvcc on this machine: a prebuilt .so is fine, none is fatal
        if not os.path.exists(path):
            raise RuntimeError(f"libsanerf_b200.so is not built and cannot be built here ({e}); "
                               "run `python -m sanerf_hq_b200.build` where nvcc is available") from e
    L = ctypes.CDLL(path)
    for name, argtypes in PROTOTYPES.items():
        if not hasattr(L, name):
            raise RuntimeError(f"{path} does not export {name}; rebuild the library")
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = (ctypes.c_char_p if name == "sanerf_error_string" else
                      ctypes.c_size_t if name.endswith("_workspace_bytes") else ctypes.c_int)
    _lib = L
    return L


def check(rc, what):
    """Non-zero status -> RuntimeError (the reference's pybind layer raises RuntimeError too)."""
    if rc != 0:
        msg = load().sanerf_error_string(int(rc)).decode()
        raise RuntimeError(f"{what}: {msg} (code {rc})")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors, what="sanerf_hq_b200"):
    """CHECK_CUDA / CHECK_CONTIGUOUS of the reference (e.g. gridencoder.cu:468-477)."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{what}: tensor must be a CUDA tensor (there is no CPU path)")
        if not t.is_contiguous():
            raise RuntimeError(f"{what}: tensor must be contiguous")


launch_counter = {"n": 0}

# Optional per-launch timing (bench.py's per-kernel roofline): when `kernel_events` is a list, every C-ABI launch made through
# `timed(label)` appends (label, start_event, end_event) recorded on the launching stream.  None (default): no events.
kernel_events = None


class timed:
    def __init__(self, label):
        self.label = label

    def __enter__(self):
        if kernel_events is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if kernel_events is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            kernel_events.append((self.label, self.a, b))
        return False


def count_launch(n=1):
    launch_counter["n"] += n
