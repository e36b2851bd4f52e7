"""GPU: the tensor-core object head (`sanerf_mask_mlp`, csrc/heads.cu) through the C ABI vs a float64 evaluation of the
reference's `mask_mlp` + compositing (nerf/renderer.py:376-385, nerf/network.py:31-66: bias-free SkipConnMLP
143 -> 256 -> 256 -> n_inst, leaky_relu(0.01) between layers; logits = sum_samples w * point_masks)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tile_transpose(x):
    """[n_rays, 32, K] -> the [ceil(n_rays*32/128)][K][128] layout sanerf_render writes for the object head."""
    n, s, k = x.shape
    rows = x.reshape(n * s, k)
    pad = (-rows.shape[0]) % 128
    if pad:
        rows = torch.cat([rows, torch.full((pad, k), float("nan"))])      # unused rows: garbage must not leak into real rows
    return rows.view(-1, 128, k).permute(0, 2, 1).contiguous()


@pytest.mark.parametrize("n_rays,n_inst", [(1, 2), (3, 2), (4, 2), (5, 3), (640, 2), (4099, 16), (20000, 2)])
def test_mask_head_matches_fp64(n_rays, n_inst):
    """records (point, geo_feat) -> in-kernel m_grid gather + mask_mlp + compositing, vs the C oracle's grid encoder
    (bit-exact restatement of the reference kernel) followed by a float64 MLP."""
    from oracle import kernels as K
    from helpers import O
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.encoders import GridEncoder
    from sanerf_hq_b200.renderer import NeRFRenderer
    lib = _lib.load()
    g = torch.Generator().manual_seed(n_rays * 31 + n_inst)
    enc = GridEncoder(num_levels=16, level_dim=8, log2_hashmap_size=19, desired_resolution=512).to(DEV)
    emb = torch.rand(enc.embeddings.shape, generator=g) * 2 - 1
    enc.embeddings.data.copy_(emb)
    sp = O.grid_spec(16, 8, 16, 19, 512)
    x01 = torch.rand(n_rays, 32, 3, generator=g)
    x01[0, 0] = torch.tensor([1.25, 0.5, 0.5])                    # outside [0,1]^3 -> zero features
    x01[0, 1] = torch.tensor([0.0, 1.0, 0.5])                     # on the boundary -> inside
    geo = (torch.rand(n_rays, 32, 15, generator=g) * 2 - 1) * 3.0
    w = torch.rand(n_rays, 32, generator=g) ** 4
    w = w / w.sum(-1, keepdim=True).clamp(min=1e-6)
    bound = lambda fan_in: 1 / fan_in ** 0.5
    ws = [(torch.rand(256, 143, generator=g) * 2 - 1) * bound(143), (torch.rand(256, 256, generator=g) * 2 - 1) * bound(256),
          (torch.rand(n_inst, 256, generator=g) * 2 - 1) * bound(256)]
    B = n_rays * 32
    feats = K.grid_encode_forward(x01.reshape(B, 3).contiguous(), emb, sp.offsets, B, 3, 8, 16, 16, np.log2(sp.per_level_scale), 16)
    feats = feats.permute(1, 0, 2).reshape(n_rays, 32, 128)
    h = torch.cat([feats, geo], dim=-1).double()
    for i, m in enumerate(ws):
        h = h @ m.double().t()
        if i < 2:
            h = torch.nn.functional.leaky_relu(h, 0.01)
    want = (w.double().unsqueeze(-1) * h).sum(-2)

    rec = _tile_transpose(torch.cat([x01, geo], dim=-1)).to(DEV)
    grid_t = _lib.GridT()
    NeRFRenderer._fill_grid(grid_t, enc)
    wd = [m.to(DEV).contiguous() for m in ws]
    wts = w.to(DEV)
    work = torch.empty(lib.sanerf_mask_head_workspace_bytes(), dtype=torch.uint8, device=DEV)
    out = torch.full((n_rays, n_inst), float("nan"), device=DEV)
    _lib.check(lib.sanerf_mask_head(_lib.ptr(rec), _lib.ptr(wts), ctypes.byref(grid_t), _lib.ptr(wd[0]), _lib.ptr(wd[1]), _lib.ptr(wd[2]),
                                    n_inst, n_rays, _lib.ptr(work), _lib.ptr(out), _lib.stream_ptr()), "mask_head")
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    # bf16 hi/lo split operands: 2^-17 per operand, random accumulation over <= 256 terms and three layers
    scale = float(want.abs().max())
    assert ((got - want).abs().max() / scale).item() < 1e-4
    assert rel_err(got, want, floor=0.1 * float(want.pow(2).mean().sqrt())) < 1e-3      # the render-path tolerance for logits
    # same result when the rays are processed with a different tiling (pointer shifted by one tile = 4 rays)
    if n_rays > 8:
        out2 = torch.empty(n_rays - 4, n_inst, device=DEV)
        _lib.check(lib.sanerf_mask_head(rec[1:].contiguous().data_ptr(), wts[4:].contiguous().data_ptr(), ctypes.byref(grid_t), _lib.ptr(wd[0]),
                                        _lib.ptr(wd[1]), _lib.ptr(wd[2]), n_inst, n_rays - 4, _lib.ptr(work), _lib.ptr(out2),
                                        _lib.stream_ptr()), "mask_head")
        assert torch.equal(out2, out[4:])


def test_mask_head_rejects_bad_arguments():
    """Argument validation of `sanerf_mask_head` (status codes instead of the reference's unchecked launches): misaligned records
    (the geo_feat rows travel by TMA bulk copy: 16-byte alignment), more instances than the padded last layer, wrong grid."""
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.encoders import GridEncoder
    from sanerf_hq_b200.renderer import NeRFRenderer
    lib = _lib.load()
    enc = GridEncoder(num_levels=16, level_dim=8, log2_hashmap_size=19, desired_resolution=512).to(DEV)
    grid_t = _lib.GridT()
    NeRFRenderer._fill_grid(grid_t, enc)
    rec = torch.zeros(2, 18, 128, device=DEV)
    wts = torch.zeros(4, 32, device=DEV)
    w0, w1, w2 = torch.zeros(256, 143, device=DEV), torch.zeros(256, 256, device=DEV), torch.zeros(2, 256, device=DEV)
    work = torch.empty(lib.sanerf_mask_head_workspace_bytes(), dtype=torch.uint8, device=DEV)
    out = torch.zeros(4, 2, device=DEV)

    def call(rec_ptr, n_inst, grid):
        return lib.sanerf_mask_head(ctypes.c_void_p(rec_ptr), _lib.ptr(wts), ctypes.byref(grid), _lib.ptr(w0), _lib.ptr(w1), _lib.ptr(w2),
                                    n_inst, 4, _lib.ptr(work), _lib.ptr(out), _lib.stream_ptr())

    assert call(rec.data_ptr(), 2, grid_t) == 0
    assert call(rec.data_ptr() + 4, 2, grid_t) != 0                  # misaligned records
    assert call(rec.data_ptr(), 17, grid_t) != 0                     # n_inst > 16
    enc4 = GridEncoder(num_levels=16, level_dim=2, log2_hashmap_size=19, desired_resolution=512).to(DEV)
    grid4 = _lib.GridT()
    NeRFRenderer._fill_grid(grid4, enc4)
    assert call(rec.data_ptr(), 2, grid4) != 0                       # the object grid has 8 channels per level
    with pytest.raises(RuntimeError):
        _lib.check(call(rec.data_ptr() + 4, 2, grid_t), "mask_head")
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("n_rays", [1, 127, 128, 129, 4000, 70001])
def test_samvit_head_matches_fp64(n_rays):
    """sanerf_samvit_mlp vs float64: five biased layers, skip concat (hidden first) before layer 2, leaky_relu, LayerNorm."""
    import ctypes
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(n_rays)
    x = torch.randn(n_rays, 163, generator=g)
    dims = [(256, 163), (256, 256), (256, 419), (256, 256), (256, 256)]
    ws = [(torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5 for o, i in dims]
    bs = [(torch.rand(o, generator=g) * 2 - 1) / i ** 0.5 for o, i in dims]
    ln_w, ln_b = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.1
    h = x.double()
    for l in range(5):
        if l == 2:
            h = torch.cat([h, x.double()], dim=-1)
        h = h @ ws[l].double().t() + bs[l].double()
        if l != 4:
            h = torch.nn.functional.leaky_relu(h, 0.01)
    want = torch.nn.functional.layer_norm(h, (256,), ln_w.double(), ln_b.double(), 1e-5)

    pad = -(-n_rays // 128) * 128
    xin = torch.full((pad, 163), float("nan"), device=DEV)
    xin[:n_rays] = x.to(DEV)
    wd, bd = [w.to(DEV).contiguous() for w in ws], [b.to(DEV).contiguous() for b in bs]
    lw, lb = ln_w.to(DEV), ln_b.to(DEV)
    work = torch.empty(lib.sanerf_samvit_mlp_workspace_bytes(), dtype=torch.uint8, device=DEV)
    out = torch.full((n_rays, 256), float("nan"), device=DEV)
    wp = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in wd])
    bp = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in bd])
    _lib.check(lib.sanerf_samvit_mlp(_lib.ptr(xin), wp, bp, _lib.ptr(lw), _lib.ptr(lb), n_rays, _lib.ptr(work), _lib.ptr(out),
                                     _lib.stream_ptr()), "samvit_mlp")
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 2e-4 * float(want.abs().max())
    assert rel_err(got, want, floor=0.1 * float(want.pow(2).mean().sqrt())) < 1e-3


def test_feature_frame_nchw_and_bilinear_resize_match_the_reference_consumer():
    """`samvit.reshape(1,h,w,256).permute(0,3,1,2).contiguous()` + `F.interpolate(..., mode='bilinear')`
    (nerf/trainer.py:540-546) produced by the head's epilogue (`feature_layout="nchw"`) and, when the size changes, one fused
    permute + resize pass (`sanerf_feature_resize_nchw`)."""
    import torch.nn.functional as F
    from helpers import build_model, frame_rays, make_case
    opt, params, specs = make_case(with_sam=True)
    model = build_model(opt, params)
    h, w = 24, 40
    rays_o, rays_d = frame_rays(800, 800, pose_k=2, rows=(300, 300 + h), cols=(100, 100 + w))
    rays_o, rays_d = rays_o.to(DEV), rays_d.to(DEV)
    with torch.no_grad():
        base = model.render(rays_o, rays_d, staged=False, perturb=False, return_feats=1, H=h, W=w)
        nhwc = base["samvit"]
        want = nhwc.reshape(1, h, w, 256).permute(0, 3, 1, 2).contiguous()
        same = model.render(rays_o, rays_d, staged=False, perturb=False, return_feats=1, H=h, W=w, feature_layout="nchw")
        assert "samvit" not in same and same["samvit_nchw"].shape == (1, 256, h, w)
        assert torch.equal(same["samvit_nchw"], want)                              # the epilogue only changes where it stores
        assert torch.equal(F.interpolate(want, (h, w), mode="bilinear"), want)     # same-size bilinear == identity
        for size in ((64, 64), (7, 13), (48, 80)):                                 # up- and down-sampling, non-integer ratios
            got = model.render(rays_o, rays_d, staged=False, perturb=False, return_feats=1, H=h, W=w, feature_layout="nchw",
                               feature_size=size)["samvit_nchw"]
            ref = F.interpolate(want, size, mode="bilinear")
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) <= 2e-6 * float(ref.abs().max()), size
        # the raw operator on a random NHWC tensor with C not a multiple of 256
        from sanerf_hq_b200 import _lib
        x = torch.randn(17, 29, 300, device=DEV)
        out = torch.empty(300, 64, 64, device=DEV)
        _lib.check(_lib.load().sanerf_feature_resize_nchw(_lib.ptr(x), 17, 29, 300, 64, 64, _lib.ptr(out), _lib.stream_ptr()), "resize")
        ref = F.interpolate(x.permute(2, 0, 1)[None].contiguous(), (64, 64), mode="bilinear")[0]
        assert float((out - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
