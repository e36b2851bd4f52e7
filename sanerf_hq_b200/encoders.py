"""Encoder nn.Modules on top of libsanerf_b200 -- drop-in for the reference's Python classes.

Mirrors (same constructor arguments, attributes, forward signatures, error behaviour):
  GridEncoder   gridencoder/grid.py:102-204    (+ autograd Function `_grid_encode`, :24-95)
  SHEncoder     shencoder/sphere_harmonics.py:61-90  (+ `_sh_encoder`, :14-53)
  FreqEncoder   freqencoder/freq.py:55-76      (+ `_freq_encoder`, :15-50)
The reference's pybind `_backend` calls are replaced by ctypes calls into the C ABI
(include/sanerf_b200.h) on the current CUDA stream.  There is no CPU path.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


def _fwd(fn):
    return torch.amp.custom_fwd(fn, device_type="cuda", cast_inputs=torch.float32)


def _bwd(fn):
    return torch.amp.custom_bwd(fn, device_type="cuda")


class _grid_encode(Function):
    """Same positional signature as the reference Function (grid.py:27) plus a trailing `bound`:
    bound > 0 means `inputs` are raw positions in [-bound, bound] and the [0,1] mapping of
    grid.py:156 is applied inside the kernel."""

    @staticmethod
    @_fwd
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0, max_level=None, bound=0.0):
        inputs = inputs.contiguous()
        embeddings = embeddings.contiguous()
        _lib.require_cuda(inputs, embeddings, offsets, what="grid_encode_forward")
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        max_level = L if max_level is None else min(max_level, L)
        lib = _lib.load()
        with torch.cuda.device(inputs.device):
            st = _lib.stream_ptr()
            if calc_grad_inputs:
                # reference layout: [L,B,C] + dy_dx, inputs must already be in [0,1]
                assert bound == 0.0
                alloc = torch.zeros if max_level < L else torch.empty
                out = alloc(L, B, C, device=inputs.device, dtype=torch.float32)
                dy_dx = alloc(B, L * D * C, device=inputs.device, dtype=torch.float32)
                _lib.check(lib.sanerf_grid_encode_forward(_lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                                                          _lib.ptr(out), B, D, C, L, max_level, S, H, _lib.ptr(dy_dx),
                                                          gridtype, int(align_corners), interpolation, st),
                           "grid_encode_forward")
                outputs = out.permute(1, 0, 2).reshape(B, L * C)
            else:
                dy_dx = None
                alloc = torch.zeros if max_level < L else torch.empty
                outputs = alloc(B, L * C, device=inputs.device, dtype=torch.float32)
                _lib.check(lib.sanerf_grid_encode_forward_fused(_lib.ptr(inputs), float(bound), _lib.ptr(embeddings),
                                                                _lib.ptr(offsets), _lib.ptr(outputs), B, D, C, L,
                                                                max_level, S, H, gridtype, int(align_corners),
                                                                interpolation, st), "grid_encode_forward_fused")
            _lib.count_launch()
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation, max_level, float(bound)]
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    @_bwd
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, max_level, bound = ctx.dims
        lib = _lib.load()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = None
        with torch.cuda.device(inputs.device):
            st = _lib.stream_ptr()
            if dy_dx is not None:
                g = grad.view(B, L, C).permute(1, 0, 2).contiguous().float()
                grad_inputs = torch.zeros_like(inputs)
                _lib.check(lib.sanerf_grid_encode_backward(_lib.ptr(g), _lib.ptr(inputs), _lib.ptr(embeddings),
                                                           _lib.ptr(offsets), _lib.ptr(grad_embeddings), B, D, C, L,
                                                           max_level, S, H, _lib.ptr(dy_dx), _lib.ptr(grad_inputs),
                                                           gridtype, int(ctx.align_corners), interpolation, st),
                           "grid_encode_backward")
                _lib.count_launch(2)
            else:
                g = grad.contiguous().float()
                _lib.check(lib.sanerf_grid_encode_backward_fused(_lib.ptr(g), _lib.ptr(inputs), bound, _lib.ptr(offsets),
                                                                 _lib.ptr(grad_embeddings), B, D, C, L, max_level, S, H,
                                                                 gridtype, int(ctx.align_corners), interpolation, st),
                           "grid_encode_backward_fused")
                _lib.count_launch()
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    """Multiresolution hash / tiled grid (reference gridencoder/grid.py:102-204)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:  # overrides per_level_scale (grid.py:107-108)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        # row offsets per level (grid.py:124-135): float64 resolution, capped at 2^T, rounded up to x8
        self.max_params = 2 ** log2_hashmap_size
        offsets, total = [], 0
        for level in range(num_levels):
            res = int(np.ceil(base_resolution * per_level_scale ** level))
            rows = int(np.ceil(min(self.max_params, res ** input_dim) / 8) * 8)
            offsets.append(total)
            total += rows
        offsets.append(total)
        self.offsets_host = list(offsets)  # python copy of the table (no device sync when packing the fused model)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(total, level_dim))
        self.reset_parameters()
        self._res_cache = None

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        top = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {top} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    def level_resolutions(self):
        """Kernel-side per-level resolutions, evaluated ON THE DEVICE with the reference kernel's fp32
        recipe (gridencoder.cu:133); they differ from the host-side float64 values used for `offsets`
        at some levels.  Cached; returns a python list."""
        if self._res_cache is None:
            dev = self.embeddings.device
            if dev.type != "cuda":
                raise RuntimeError("GridEncoder.level_resolutions needs the module on a CUDA device")
            out = torch.empty(self.num_levels, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.load().sanerf_grid_level_resolutions(_lib.ptr(out), self.num_levels,
                                                                    float(np.log2(self.per_level_scale)),
                                                                    int(self.base_resolution), _lib.stream_ptr()),
                           "grid_level_resolutions")
            self._res_cache = [int(v) for v in out.cpu().tolist()]
        return self._res_cache

    def forward(self, inputs, bound=1, max_level=None):
        # inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        prefix_shape = list(inputs.shape[:-1])
        if inputs.requires_grad:
            x = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            out = grid_encode(x, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, True,
                              self.gridtype_id, self.align_corners, self.interp_id, max_level, 0.0)
        else:
            out = grid_encode(inputs.reshape(-1, self.input_dim), self.embeddings, self.offsets, self.per_level_scale,
                              self.base_resolution, False, self.gridtype_id, self.align_corners, self.interp_id,
                              max_level, float(bound))
        return out.view(prefix_shape + [self.output_dim])

    @torch.amp.autocast("cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        D, C, L = self.input_dim, self.embeddings.shape[1], self.offsets.shape[0] - 1
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim).contiguous()
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        _lib.require_cuda(inputs, self.embeddings, self.embeddings.grad, what="grad_total_variation")
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().sanerf_grad_total_variation(
                _lib.ptr(inputs), _lib.ptr(self.embeddings), _lib.ptr(self.embeddings.grad), _lib.ptr(self.offsets),
                float(weight), B, D, C, L, float(np.log2(self.per_level_scale)), int(self.base_resolution),
                self.gridtype_id, int(self.align_corners), _lib.stream_ptr()), "grad_total_variation")
            _lib.count_launch()

    @torch.amp.autocast("cuda", enabled=False)
    def grad_weight_decay(self, weight=0.1):
        B, C, L = self.embeddings.shape[0], self.embeddings.shape[1], self.offsets.shape[0] - 1
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        _lib.require_cuda(self.embeddings, self.embeddings.grad, what="grad_weight_decay")
        with torch.cuda.device(self.embeddings.device):
            _lib.check(_lib.load().sanerf_grad_weight_decay(_lib.ptr(self.embeddings), _lib.ptr(self.embeddings.grad),
                                                           _lib.ptr(self.offsets), float(weight), B, C, L,
                                                           _lib.stream_ptr()), "grad_weight_decay")
            _lib.count_launch()


class _sh_encoder(Function):
    @staticmethod
    @_fwd
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous()
        _lib.require_cuda(inputs, what="sh_encode_forward")
        B, input_dim = inputs.shape
        out_dim = degree ** 2
        outputs = torch.empty(B, out_dim, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * out_dim, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().sanerf_sh_encode_forward(_lib.ptr(inputs), _lib.ptr(outputs), B, input_dim, degree,
                                                           _lib.ptr(dy_dx), _lib.stream_ptr()), "sh_encode_forward")
            _lib.count_launch()
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = [B, input_dim, degree]
        return outputs

    @staticmethod
    @_bwd
    def backward(ctx, grad):
        inputs, dy_dx = ctx.saved_tensors
        if dy_dx is None:
            return None, None, None
        grad = grad.contiguous()
        B, input_dim, degree = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().sanerf_sh_encode_backward(_lib.ptr(grad), _lib.ptr(inputs), B, input_dim, degree,
                                                            _lib.ptr(dy_dx), _lib.ptr(grad_inputs), _lib.stream_ptr()),
                       "sh_encode_backward")
            _lib.count_launch()
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    """Real spherical harmonics of a direction (reference shencoder/sphere_harmonics.py:61-90)."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        inputs = inputs / torch.norm(inputs, dim=-1, keepdim=True)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])


class _freq_encoder(Function):
    @staticmethod
    @_fwd
    def forward(ctx, inputs, degree, output_dim):
        if not inputs.is_cuda:
            inputs = inputs.cuda()  # freq.py:22
        inputs = inputs.contiguous()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().sanerf_freq_encode_forward(_lib.ptr(inputs), B, input_dim, degree, output_dim,
                                                             _lib.ptr(outputs), _lib.stream_ptr()), "freq_encode_forward")
            _lib.count_launch()
        ctx.save_for_backward(inputs, outputs)
        ctx.dims = [B, input_dim, degree, output_dim]
        return outputs

    @staticmethod
    @_bwd
    def backward(ctx, grad):
        grad = grad.contiguous()
        inputs, outputs = ctx.saved_tensors
        B, input_dim, degree, output_dim = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().sanerf_freq_encode_backward(_lib.ptr(grad), _lib.ptr(outputs), B, input_dim, degree,
                                                              output_dim, _lib.ptr(grad_inputs), _lib.stream_ptr()),
                       "freq_encode_backward")
            _lib.count_launch()
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    """NeRF sinusoidal encoding (reference freqencoder/freq.py:55-76)."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = freq_encode(inputs, self.degree, self.output_dim)
        return outputs.reshape(prefix_shape + [self.output_dim])
