"""trunc_exp -- density activation (reference activation.py:5-17).

forward: exp(x) evaluated in fp32; backward: g * exp(clamp(x, -15, 15)) so a huge pre-activation
cannot blow up the gradient.
"""
import torch
from torch.autograd import Function


class _trunc_exp(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply
