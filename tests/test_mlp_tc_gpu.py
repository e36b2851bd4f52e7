"""GPU: the tcgen05 (tensor-core, split-precision 3xTF32) MLP path of csrc/tc.cuh through the C ABI
(`sanerf_mlp3_tc`) vs a float64 torch evaluation of the reference's `MLP` (nerf/network.py:9-29: bias-free Linear
layers, ReLU between, none after the last).  Tolerance: fp32-equivalent -- the error against float64 must be no worse
than a few times the error of a plain fp32 torch evaluation (which rounds every product to fp32 as well)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref64(x, ws):
    h = x.double()
    for i, w in enumerate(ws):
        h = h @ w.double().t()
        if i + 1 < len(ws):
            h = torch.relu(h)
    return h


@pytest.mark.parametrize("K,H", [(32, 64), (8, 16), (16, 32)])
@pytest.mark.parametrize("M", [1, 127, 128, 129, 512, 70001])
def test_mlp3_tensor_core_matches_fp64(K, H, M):
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(K * 1000 + M)
    x = (torch.rand(M, K, generator=g) * 2 - 1)
    x[0] = 0
    if M > 5:
        x[5] = 1e-3 * x[5]          # small activations: lo parts become subnormal-free but tiny
        x[3] = 50 * x[3]            # large activations
    bound = lambda fan_in: 1 / fan_in ** 0.5
    ws = [(torch.rand(H, K, generator=g) * 2 - 1) * bound(K), (torch.rand(H, H, generator=g) * 2 - 1) * bound(H),
          (torch.rand(16, H, generator=g) * 2 - 1) * bound(H)]
    xd, wd = x.to(DEV), [w.to(DEV).contiguous() for w in ws]
    out = torch.full((M, 16), float("nan"), device=DEV)
    _lib.check(lib.sanerf_mlp3_tc(_lib.ptr(xd), _lib.ptr(wd[0]), _lib.ptr(wd[1]), _lib.ptr(wd[2]), _lib.ptr(out), M, K, H,
                                  _lib.stream_ptr()), "mlp3_tc")
    torch.cuda.synchronize()
    want = _ref64(x, ws)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    scale = want.abs().max().clamp(min=1e-6)
    err_tc = ((got - want).abs().max() / scale).item()
    fp32 = x
    for i, w in enumerate(ws):
        fp32 = fp32 @ w.t()
        if i < 2:
            fp32 = torch.relu(fp32)
    err_fp32 = ((fp32.double() - want).abs().max() / scale).item()
    # 3xTF32 drops the lo*lo term (2^-22 per product) and the tensor core aligns the K=8 products before adding: allow
    # 16x the fp32 evaluation's own error, floor 4e-6 of the output scale (three layers, up to 64 terms each)
    assert err_tc <= max(16 * err_fp32, 4e-6), (err_tc, err_fp32)
    assert rel_err(got, want, floor=float(scale) * 1e-2) < 2e-5
