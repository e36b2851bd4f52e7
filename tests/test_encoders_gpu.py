"""GPU: the encoder operators of libsanerf_b200 (through the C ABI, via the drop-in nn.Modules) vs the CPU
oracle (oracle/sanerf_oracle.c), on seeded inputs.  Bit-exact where the arithmetic is the same FMA
sequence (grid forward), tight fp32 tolerance elsewhere (stated per test)."""

import numpy as np
import pytest
import torch

from helpers import O, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(B, seed=0, oob=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, generator=g)
    x[0] = 0.0
    x[1] = 1.0
    x[2] = torch.tensor([0.5, 0.5, 0.5])
    if oob:
        x[3] = torch.tensor([1.25, 0.5, 0.5])
        x[4] = torch.tensor([0.5, -0.01, 0.5])
    # points exactly on cell boundaries of several levels
    x[5] = torch.tensor([1 / 16 + 0.5 / 16, 3 / 32, 0.75])
    return x


GRIDS = [  # (L, C, log2T, desired)  -- config #1, config #2 grid, proposal grids, feature grid
    (4, 2, 19, 4096), (16, 2, 19, 4096), (5, 2, 17, 128), (5, 2, 17, 256), (16, 8, 19, 512), (4, 8, 19, 512),
    (8, 1, 14, 512), (6, 4, 15, 1024)]


@pytest.mark.parametrize("L,C,T,des", GRIDS)
def test_grid_forward_matches_oracle_bit_exact(L, C, T, des):
    from oracle import kernels as K
    from sanerf_hq_b200.encoders import GridEncoder
    enc = GridEncoder(num_levels=L, level_dim=C, log2_hashmap_size=T, desired_resolution=des).to(DEV)
    g = torch.Generator().manual_seed(L * 100 + C)
    emb = torch.rand(enc.embeddings.shape, generator=g) * 2 - 1
    enc.embeddings.data.copy_(emb)
    sp = O.grid_spec(L, C, 16, T, des)
    assert torch.equal(enc.offsets.cpu(), sp.offsets)
    # device-evaluated level resolutions == the oracle's (glibc exp2f) table
    S = np.log2(sp.per_level_scale)
    assert enc.level_resolutions() == [K.level_resolution(l, S, 16) for l in range(L)]

    B = 20000
    x01 = _inputs(B, seed=C)
    want = K.grid_encode_forward(x01, emb, sp.offsets, B, 3, C, L, L, S, 16)          # [L,B,C]
    want_blc = want.permute(1, 0, 2).reshape(B, L * C)
    with torch.no_grad():
        # module path: raw positions in [-bound, bound]; x01 = (x + 2) / 4 must round-trip exactly -> use x = 4*x01 - 2
        xs = (x01 * 4 - 2)
        ok = ((xs + 2) / 4 == x01).all(dim=-1)
        got = enc(xs.to(DEV), bound=2).cpu()
    assert torch.equal(got[ok], want_blc[ok]), "fused [B,L*C] kernel differs from oracle"
    # reference-layout entry point [L,B,C] straight through the C ABI
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    out = torch.full((L, B, C), 7.0, device=DEV)
    xd = x01.to(DEV)
    _lib.check(lib.sanerf_grid_encode_forward(_lib.ptr(xd), _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets), _lib.ptr(out), B, 3, C, L, L,
                                              float(S), 16, None, 0, 0, 0, _lib.stream_ptr()), "fwd")
    assert torch.equal(out.cpu(), want)
    # max_level < L writes only the first levels
    out2 = torch.zeros((L, B, C), device=DEV)
    _lib.check(lib.sanerf_grid_encode_forward(_lib.ptr(xd), _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets), _lib.ptr(out2), B, 3, C, L, 2,
                                              float(S), 16, None, 0, 0, 0, _lib.stream_ptr()), "fwd")
    assert torch.equal(out2.cpu()[:2], want[:2]) and torch.all(out2[2:] == 0)


@pytest.mark.parametrize("gridtype,align,interp", [(1, False, 0), (0, True, 0), (0, False, 1), (1, True, 1)])
def test_grid_forward_variants_and_dy_dx(gridtype, align, interp):
    """tiled grid, align_corners, smoothstep + the analytic dy_dx output (2-D and 3-D)."""
    from oracle import kernels as K
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    for D in (2, 3):
        L, C, H, T = 6, 2, 8, 12
        pls = 1.5
        off = K.grid_offsets(D, L, H, pls, T)
        g = torch.Generator().manual_seed(D)
        emb = torch.rand(int(off[-1]), C, generator=g) * 2 - 1
        B = 5000
        x = torch.rand(B, D, generator=g)
        S = np.log2(pls)
        dy_want = torch.zeros(B, L * D * C)
        want = K.grid_encode_forward(x, emb, off, B, D, C, L, L, S, H, dy_want, gridtype, align, interp)
        out = torch.empty(L, B, C, device=DEV)
        dy = torch.empty(B, L * D * C, device=DEV)
        xd, ed, od = x.to(DEV), emb.to(DEV), off.to(DEV)
        _lib.check(lib.sanerf_grid_encode_forward(_lib.ptr(xd), _lib.ptr(ed), _lib.ptr(od), _lib.ptr(out), B, D, C, L, L, float(S), H,
                                                  _lib.ptr(dy), gridtype, int(align), interp, _lib.stream_ptr()), "fwd")
        assert rel_err(out, want, floor=1e-2) < 1e-5
        assert rel_err(dy, dy_want, floor=1.0) < 1e-4


def test_grid_backward_tv_wd_match_oracle():
    from oracle import kernels as K
    from sanerf_hq_b200.encoders import GridEncoder
    L, C, T, des = 8, 2, 15, 512
    enc = GridEncoder(num_levels=L, level_dim=C, log2_hashmap_size=T, desired_resolution=des).to(DEV)
    g = torch.Generator().manual_seed(11)
    emb = torch.rand(enc.embeddings.shape, generator=g) * 2 - 1
    enc.embeddings.data.copy_(emb)
    sp = O.grid_spec(L, C, 16, T, des)
    S = np.log2(sp.per_level_scale)
    B = 6000
    x01 = _inputs(B, seed=5)
    xs = (x01 * 2 - 1)
    gout = torch.randn(B, L * C, generator=g)
    # autograd through the module (fused backward kernel, vector red.add)
    y = enc(xs.to(DEV), bound=1)
    y.backward(gout.to(DEV))
    x01_dev = ((xs + 1) / 2)
    want = torch.zeros_like(emb)
    K.grid_encode_backward(gout.view(B, L, C).permute(1, 0, 2).contiguous(), x01_dev, emb, sp.offsets, want, B, 3, C, L, L, S, 16)
    # atomics accumulate in arbitrary order -> fp32 summation-order tolerance
    assert rel_err(enc.embeddings.grad, want, floor=1e-1) < 1e-4
    # TV + WD regularisers, in place on .grad
    g0 = enc.embeddings.grad.clone()
    pts = torch.rand(3000, 3, generator=g)
    enc.grad_total_variation(weight=0.3, inputs=(pts * 2 - 1).to(DEV), bound=1)
    want_tv = g0.cpu().clone()
    K.grad_total_variation((pts * 2 - 1 + 1) / 2, emb, want_tv, sp.offsets, 0.3, 3000, 3, C, L, S, 16)
    assert rel_err(enc.embeddings.grad, want_tv, floor=1e-1) < 1e-4
    g1 = enc.embeddings.grad.clone()
    enc.grad_weight_decay(weight=0.25)
    want_wd = g1.cpu().clone()
    K.grad_weight_decay(emb, want_wd, sp.offsets, 0.25, emb.shape[0], C, L)
    assert rel_err(enc.embeddings.grad, want_wd, floor=1e-1) < 1e-5
    # inputs.requires_grad -> dy_dx path + input gradient (K3) against autograd-free finite differences of the oracle
    xs2 = (torch.rand(64, 3, generator=g) * 1.6 - 0.8).to(DEV).requires_grad_(True)
    y2 = enc(xs2, bound=1)
    w = torch.randn(y2.shape, generator=g).to(DEV)
    (y2 * w).sum().backward()
    # exact check: grad_inputs[b,d] = sum_{l,c} grad[b,l,c] * dy_dx[b,l,d,c] with the oracle's analytic dy_dx
    # (gridencoder.cu:352-378), then the chain rule of the (x+bound)/(2*bound) mapping (grid.py:156) = 1/2
    x01_2 = (xs2.detach().cpu() + 1) / 2
    dy_want = torch.zeros(64, L * 3 * C)
    K.grid_encode_forward(x01_2, emb, sp.offsets, 64, 3, C, L, L, S, 16, dy_want)
    want_gi = (w.cpu().view(64, L, 1, C) * dy_want.view(64, L, 3, C)).sum(dim=(1, 3)) / 2
    assert rel_err(xs2.grad, want_gi, floor=1.0) < 1e-4
    # sanity: central finite differences of the oracle agree wherever the stencil stays inside one cell of every level
    eps = 1e-4
    num = torch.zeros(64, 3)
    for d in range(3):
        xp, xm = xs2.detach().cpu().clone(), xs2.detach().cpu().clone()
        xp[:, d] += eps
        xm[:, d] -= eps
        fp = K.grid_encoder_apply(xp, emb, sp.offsets, sp.per_level_scale, 16, bound=1)
        fm = K.grid_encoder_apply(xm, emb, sp.offsets, sp.per_level_scale, 16, bound=1)
        num[:, d] = ((fp - fm) * w.cpu()).sum(-1) / (2 * eps)
    close = ((xs2.grad.cpu() - num).abs() <= 2e-2 * num.abs().clamp(min=1.0))
    assert close.float().mean() > 0.75   # ~10 % of the stencils straddle a cell face of some level at eps=1e-4


def test_unsupported_shapes_raise_like_the_reference():
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(8, 8, device=DEV)
    off = torch.tensor([0, 8], dtype=torch.int32, device=DEV)
    rc = lib.sanerf_grid_encode_forward(_lib.ptr(t), _lib.ptr(t), _lib.ptr(off), _lib.ptr(t), 1, 3, 3, 1, 1, 1.0, 16, None, 0, 0, 0, None)
    assert rc == -3                                                # C=3 -> "C must be 1, 2, 4, 8, 16 or 32" (gridencoder.cu:392)
    rc = lib.sanerf_grid_encode_forward(_lib.ptr(t), _lib.ptr(t), _lib.ptr(off), _lib.ptr(t), 1, 6, 2, 1, 1, 1.0, 16, None, 0, 0, 0, None)
    assert rc == -2                                                # D=6 (gridencoder.cu:409)
    with pytest.raises(RuntimeError):
        _lib.check(rc, "grid_encode_forward")
    assert lib.sanerf_sh_encode_forward(_lib.ptr(t), _lib.ptr(t), 1, 3, 9, None, None) == -4
    assert lib.sanerf_grid_encode_forward(None, _lib.ptr(t), _lib.ptr(off), _lib.ptr(t), 1, 3, 2, 1, 1, 1.0, 16, None, 0, 0, 0, None) == -1
    assert lib.sanerf_grid_encode_forward(None, None, None, None, 0, 3, 2, 1, 1, 1.0, 16, None, 0, 0, 0, None) == 0   # empty batch


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_matches_oracle(degree):
    from oracle import kernels as K
    from sanerf_hq_b200.encoders import SHEncoder
    g = torch.Generator().manual_seed(degree)
    d = torch.randn(4000, 3, generator=g)
    d[0] = torch.tensor([0.0, 0.0, 1.0])
    d[1] = torch.tensor([1.0, 0.0, 0.0])
    want = K.sh_encoder_apply(d, degree=degree)
    enc = SHEncoder(degree=degree)
    got = enc(d.to(DEV))
    assert got.shape == (4000, degree * degree)
    # same polynomials; nvcc contracts mul+add into FMA where the C oracle (-ffp-contract=off) rounds twice, and the
    # degree-7/8 terms are sums of up to 5 products of magnitude ~5 that cancel: 1e-5 absolute, 2e-6 relative
    assert (got.cpu() - want).abs().max().item() < 1e-5
    assert rel_err(got, want, floor=1.0) < 2e-6 * degree
    # analytic gradient (dual numbers) vs central differences of the oracle in float64-ish steps
    dd = d[:200].clone().to(DEV).requires_grad_(True)
    w = torch.randn(200, degree * degree, generator=g).to(DEV)
    (enc(dd) * w).sum().backward()
    eps = 1e-3
    num = torch.zeros(200, 3)
    for a in range(3):
        dp, dm = d[:200].clone(), d[:200].clone()
        dp[:, a] += eps
        dm[:, a] -= eps
        num[:, a] = ((K.sh_encoder_apply(dp, degree) - K.sh_encoder_apply(dm, degree)) * w.cpu()).sum(-1) / (2 * eps)
    assert rel_err(dd.grad, num, floor=1.0) < 2e-2


def test_freq_matches_oracle_and_torch_encoder():
    from oracle import kernels as K
    from sanerf_hq_b200.encoders import FreqEncoder
    from sanerf_hq_b200.encoding import get_encoder
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3000, 3, generator=g) * 2 - 1
    enc = FreqEncoder(input_dim=3, degree=6)
    got = enc(x.to(DEV)).cpu()
    want = K.freq_encode_forward(x, 3000, 3, 6, 39)
    # device uses the SFU fast sine (as the reference, -use_fast_math): abs error grows with |2^f x| <= 32
    assert (got - want).abs().max().item() < 2e-4
    ref_t, _ = get_encoder("frequency_torch", multires=6)
    assert (got - ref_t(x)).abs().max().item() < 2e-4
    xg = x[:100].clone().to(DEV).requires_grad_(True)
    w = torch.randn(100, 39, generator=g)
    (enc(xg) * w.to(DEV)).sum().backward()
    xt = x[:100].clone().requires_grad_(True)
    (ref_t(xt) * w).sum().backward()
    assert rel_err(xg.grad, xt.grad, floor=1.0) < 5e-3
