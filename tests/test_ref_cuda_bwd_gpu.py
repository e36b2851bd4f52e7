"""GPU: the regulariser and backward kernels (K4 TV, K5 weight decay, K7 SH backward, K9 freq backward) against the REFERENCE'S
OWN compiled kernels in oracle/_ref (tests/test_ref_cuda_gpu.py does the same for the grid forward / backward and the SH /
freq forward).  The same kernels are also checked against the C oracle (`test_grid_backward_tv_wd_match_oracle`,
`test_sh_matches_oracle`, `test_freq_matches_oracle_and_torch_encoder`)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu
DEV = "cuda"
REFDIR = os.path.join(REPO, "oracle", "_ref")


def _ref(name):
    if not os.path.exists(os.path.join(REFDIR, name + ".so")):
        pytest.skip(f"oracle/_ref/{name}.so not built (python oracle/build_ref.py)")
    if REFDIR not in sys.path:
        sys.path.append(REFDIR)
    return importlib.import_module(name)


@pytest.mark.parametrize("L,C,T,des", [(16, 2, 19, 4096), (16, 8, 19, 512)])
def test_total_variation_and_weight_decay_vs_reference_kernels(L, C, T, des):
    """gridencoder.cu grad_total_variation / grad_weight_decay (called from grid.py:169-207): atomics on both sides, so a
    summation-order tolerance; weight decay is a plain element-wise update -> tight."""
    ref = _ref("_gridencoder")
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.encoders import GridEncoder
    lib = _lib.load()
    enc = GridEncoder(num_levels=L, level_dim=C, log2_hashmap_size=T, desired_resolution=des).to(DEV)
    torch.manual_seed(2)
    enc.embeddings.data.uniform_(-1, 1)
    emb = enc.embeddings.data
    S = float(np.log2(enc.per_level_scale))
    B = 100000
    x = torch.rand(B, 3, device=DEV)
    want = torch.zeros_like(emb)
    ref.grad_total_variation(x, emb, want, enc.offsets, 0.37, B, 3, C, L, S, 16, 0, False)
    got = torch.zeros_like(emb)
    _lib.check(lib.sanerf_grad_total_variation(_lib.ptr(x), _lib.ptr(emb), _lib.ptr(got), _lib.ptr(enc.offsets), 0.37, B, 3, C, L, S, 16, 0, 0,
                                               _lib.stream_ptr()), "tv")
    torch.cuda.synchronize()
    assert float(want.abs().sum()) > 0
    assert ((got - want).abs() <= 1e-4 * want.abs().clamp(min=float(want.abs().max()) * 1e-3)).all()
    want = torch.zeros_like(emb)
    ref.grad_weight_decay(emb, want, enc.offsets, 0.1, emb.shape[0], C, L)
    got = torch.zeros_like(emb)
    _lib.check(lib.sanerf_grad_weight_decay(_lib.ptr(emb), _lib.ptr(got), _lib.ptr(enc.offsets), 0.1, emb.shape[0], C, L, _lib.stream_ptr()), "wd")
    torch.cuda.synchronize()
    assert float(want.abs().sum()) > 0
    assert ((got - want).abs() <= 1e-6 * want.abs().clamp(min=1e-12)).all()


def test_sh_and_freq_backward_vs_reference_kernels():
    refs, reff = _ref("_shencoder"), _ref("_freqencoder")
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(4)
    B = 50000
    d = torch.nn.functional.normalize(torch.randn(B, 3, device=DEV), dim=-1)
    for deg in (4, 8):
        out = torch.empty(B, deg * deg, device=DEV)
        dy = torch.empty(B, 3 * deg * deg, device=DEV)
        refs.sh_encode_forward(d, out, B, 3, deg, dy)
        g = torch.randn(B, deg * deg, device=DEV)
        want = torch.zeros(B, 3, device=DEV)
        refs.sh_encode_backward(g, d, B, 3, deg, dy, want)
        got = torch.zeros(B, 3, device=DEV)
        _lib.check(lib.sanerf_sh_encode_backward(_lib.ptr(g), _lib.ptr(d), B, 3, deg, _lib.ptr(dy), _lib.ptr(got), _lib.stream_ptr()), "sh bwd")
        torch.cuda.synchronize()
        assert ((got - want).abs() <= 1e-5 * want.abs().clamp(min=1.0)).all()
    x = torch.rand(B, 3, device=DEV) * 2 - 1
    out = torch.empty(B, 39, device=DEV)
    reff.freq_encode_forward(x, B, 3, 6, 39, out)
    g = torch.randn(B, 39, device=DEV)
    want = torch.zeros(B, 3, device=DEV)
    reff.freq_encode_backward(g, out, B, 3, 6, 39, want)
    got = torch.zeros(B, 3, device=DEV)
    _lib.check(lib.sanerf_freq_encode_backward(_lib.ptr(g), _lib.ptr(out), B, 3, 6, 39, _lib.ptr(got), _lib.stream_ptr()), "freq bwd")
    torch.cuda.synchronize()
    assert ((got - want).abs() <= 1e-5 * want.abs().clamp(min=1.0)).all()
