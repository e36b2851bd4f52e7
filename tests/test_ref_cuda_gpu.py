"""GPU: new kernels vs the REFERENCE'S OWN CUDA kernels (oracle/_ref/*.so = the reference .cu/.cpp compiled
verbatim by oracle/build_ref.py, flag-only change -std=c++17).  Skipped when oracle/_ref was not built."""
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


@pytest.mark.parametrize("L,C,T,des", [(16, 2, 19, 4096), (5, 2, 17, 128), (16, 8, 19, 512)])
def test_grid_forward_bit_exact_vs_reference_kernel(L, C, T, des):
    ref = _ref("_gridencoder")
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.encoders import GridEncoder
    enc = GridEncoder(num_levels=L, level_dim=C, log2_hashmap_size=T, desired_resolution=des).to(DEV)
    torch.manual_seed(1)
    enc.embeddings.data.uniform_(-1, 1)
    B = 200000
    x = torch.rand(B, 3, device=DEV)
    x[:3] = torch.tensor([[0.0, 0, 0], [1.0, 1, 1], [1.5, 0.2, 0.2]], device=DEV)
    S = float(np.log2(enc.per_level_scale))
    want = torch.empty(L, B, C, device=DEV)
    ref.grid_encode_forward(x, enc.embeddings.data, enc.offsets, want, B, 3, C, L, L, S, 16, None, 0, False, 0)
    got = torch.empty(L, B, C, device=DEV)
    lib = _lib.load()
    _lib.check(lib.sanerf_grid_encode_forward(_lib.ptr(x), _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets), _lib.ptr(got), B, 3, C, L, L, S, 16,
                                              None, 0, 0, 0, _lib.stream_ptr()), "fwd")
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    # backward (atomics on both sides -> summation-order tolerance) and dy_dx
    g = torch.randn(L, B, C, device=DEV)
    ge_ref = torch.zeros_like(enc.embeddings.data)
    ref.grid_encode_backward(g, x, enc.embeddings.data, enc.offsets, ge_ref, B, 3, C, L, L, S, 16, None, None, 0, False, 0)
    ge = torch.zeros_like(ge_ref)
    _lib.check(lib.sanerf_grid_encode_backward(_lib.ptr(g), _lib.ptr(x), _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets), _lib.ptr(ge), B, 3, C, L,
                                               L, S, 16, None, None, 0, 0, 0, _lib.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    assert ((ge - ge_ref).abs() <= 1e-4 * ge_ref.abs().clamp(min=1.0)).all()


def test_sh_and_freq_vs_reference_kernels():
    ref = _ref("_shencoder")
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(100000, 3, device=DEV), dim=-1)
    for deg in (4, 8):
        want = torch.empty(100000, deg * deg, device=DEV)
        dy_want = torch.empty(100000, 3 * deg * deg, device=DEV)
        ref.sh_encode_forward(d, want, 100000, 3, deg, dy_want)
        got = torch.empty_like(want)
        dy = torch.empty_like(dy_want)
        _lib.check(lib.sanerf_sh_encode_forward(_lib.ptr(d), _lib.ptr(got), 100000, 3, deg, _lib.ptr(dy), _lib.stream_ptr()), "sh")
        torch.cuda.synchronize()
        assert (got - want).abs().max().item() < 4e-6
        assert (dy - dy_want).abs().max().item() < 1e-4
    reff = _ref("_freqencoder")
    x = torch.rand(50000, 3, device=DEV) * 2 - 1
    want = torch.empty(50000, 39, device=DEV)
    reff.freq_encode_forward(x, 50000, 3, 6, 39, want)
    got = torch.empty_like(want)
    _lib.check(lib.sanerf_freq_encode_forward(_lib.ptr(x), 50000, 3, 6, 39, _lib.ptr(got), _lib.stream_ptr()), "freq")
    torch.cuda.synchronize()
    assert (got - want).abs().max().item() < 1e-6      # same __sinf(scalbnf(x,f)+phase) expression
