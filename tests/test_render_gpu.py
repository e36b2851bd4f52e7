"""GPU parity of the render hot path: fused megakernel (sanerf_render through NeRFRenderer.render) and the
composed torch path vs (a) the golden fixtures generated from the reference itself and (b) the CPU oracle
on seeded inputs.  Tolerance: |cand - ref| <= 1e-3 * max(|ref|, 1e-3); index buffers
bit-exact except where the oracle's cdf is within a few ulps of the query."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, REL_TOL, assert_close, build_model, frame_rays, index_mismatch_report, make_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = ["cfg1_rgb", "cfg1_sam", "cfg1_mask", "full_rgb", "full_sam", "full_mask"]


def _render(model, rays_o, rays_d, staged, fused, taps=None, **kw):
    model.fused = fused
    with torch.no_grad():
        if taps is not None and fused:
            return model._run_fused(rays_o.to(DEV), rays_d.to(DEV), taps=taps, **kw)
        return model.render(rays_o.to(DEV), rays_d.to(DEV), staged=staged, **kw)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "composed"])
@pytest.mark.parametrize("name", CASES)
def test_render_matches_reference_fixture(name, fused):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    small, with_sam, with_mask, h, w, batch, staged = [int(v) for v in fx["meta"]]
    opt, params, specs = make_case(small=bool(small), with_sam=bool(with_sam), with_mask=bool(with_mask), max_ray_batch=batch)
    model = build_model(opt, params, small=bool(small))
    rays_o, rays_d = torch.from_numpy(fx["rays_o"]), torch.from_numpy(fx["rays_d"])
    kw = dict(perturb=False, bg_color=1)
    if with_sam:
        kw.update(return_feats=1, H=h, W=w)
    if with_mask:
        kw.update(return_mask=1)
    out = _render(model, rays_o, rays_d, bool(staged), fused, **kw)
    want_keys = {k[4:] for k in fx.files if k.startswith("out_")}
    assert {k for k, v in out.items() if torch.is_tensor(v)} == want_keys
    for k in sorted(want_keys):
        assert out[k].shape == fx["out_" + k].shape, k
        assert_close(out[k], fx["out_" + k], REL_TOL, f"{name}/{k}")
    if fused:
        taps = dict(inds0=None, inds1=None)
        _render(model, rays_o, rays_d, False, True, taps=taps, **{k: v for k, v in kw.items() if k not in ("return_feats", "H", "W")})
        # against the fixture's index buffers: mismatches must be rare and off by one (near-ties)
        for t, ref in (("inds0", fx["inds0"]), ("inds1", fx["inds1"])):
            got = taps[t].cpu().numpy()
            assert np.abs(got.astype(np.int32) - ref).max() <= 1
            assert (got != ref).mean() < 2e-3, t


@pytest.mark.parametrize("mode", ["rgb", "sam", "mask"])
def test_fused_vs_oracle_seeded_rays(mode):
    """Default-size network, a sub-block of the 800x800 frame (oracle finishes in seconds), every output + the
    per-stage taps; also a peakier scene (table_scale=3) and explicit cam_near_far / bg colour."""
    for table_scale, cnf, bg in ((1.0, None, None), (3.0, torch.tensor([[0.3, 6.0]]), torch.tensor([0.2, 0.5, 0.9]))):
        opt, params, specs = make_case(with_sam=mode == "sam", with_mask=mode == "mask", table_scale=table_scale)
        model = build_model(opt, params)
        rays_o, rays_d = frame_rays(800, 800, pose_k=5, rows=(200, 216), cols=(384, 416))   # 512 rays
        N = rays_o.shape[0]
        kw = {}
        if mode == "sam":
            kw.update(return_feats=1, H=16, W=32)
        if mode == "mask":
            kw.update(return_mask=1)
        ref, ex = O.run(params, specs, opt, rays_o, rays_d, bg_color=bg, cam_near_far=cnf, **kw)
        taps = dict(inds0=None, inds1=None, weights2=None, sigma2=None, bins2=None, f_image=None)
        out = _render(model, rays_o, rays_d, False, True, taps=taps, bg_color=None if bg is None else bg.to(DEV),
                      cam_near_far=None if cnf is None else cnf.to(DEV), **kw)
        for k, v in ref.items():
            assert_close(out[k], v, REL_TOL, f"{mode}/scale{table_scale}/{k}")
        assert_close(taps["bins2"], ex["bins"][2], REL_TOL, "bins2")
        assert_close(taps["weights2"], ex["weights"][2], REL_TOL, "weights2")
        assert_close(taps["f_image"], ex["f_image"], REL_TOL, "f_image")
        for i, name in enumerate(("inds0", "inds1")):
            aux = ex["pdf"][i]
            n_bad, n_unexpl = index_mismatch_report(taps[name].cpu().numpy(), aux["inds"].numpy(), aux["cdf"].numpy(), aux["u"].numpy())
            assert n_unexpl == 0, f"{name}: {n_unexpl} unexplained index mismatches of {n_bad}"
            assert n_bad <= 2e-3 * aux["inds"].numel()


def test_sample_pdf_index_buffers_bit_exact():
    """Standalone sample_pdf on identical inputs: the only integer buffers on the path.  Identical weights/bins in,
    so the index buffers must be equal to the oracle's except at ulp-level ties in the cdf."""
    from sanerf_hq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for T0, T in ((128, 65), (64, 33)):
        N = 4096
        w = torch.rand(N, T0, generator=g) ** 6
        w[:8] = 0                                  # empty rays: pdf uniform
        w[8:16, : T0 // 2] = 0
        w[16, 5] = 1e30                            # one dominant sample
        edges = torch.sort(torch.rand(N, T0 + 1, generator=g), dim=-1).values
        edges[:, 0], edges[:, -1] = 0, 1
        new_ref, aux = O.sample_pdf(edges, w, T)
        u = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T).to(DEV)
        new = torch.empty(N, T, device=DEV)
        inds = torch.empty(N, T, dtype=torch.int16, device=DEV)
        bd, wd = edges.to(DEV), w.to(DEV)
        _lib.check(lib.sanerf_sample_pdf(_lib.ptr(bd), _lib.ptr(wd), _lib.ptr(u), N, T0, T, _lib.ptr(new), _lib.ptr(inds), _lib.stream_ptr()), "pdf")
        n_bad, n_unexpl = index_mismatch_report(inds.cpu().numpy(), aux["inds"].numpy(), aux["cdf"].numpy(), aux["u"].numpy())
        assert n_unexpl == 0 and n_bad <= 1e-3 * N * T, (n_bad, n_unexpl)
        same = (inds.cpu().long() == aux["inds"]).all(dim=-1)
        assert rel_err(new.cpu()[same], new_ref[same], floor=1e-3) < 1e-4
        # properties: outputs sorted, inside [0,1]
        assert bool((new[:, 1:] >= new[:, :-1]).all()) and float(new.min()) >= 0 and float(new.max()) <= 1


@pytest.mark.parametrize("mode", ["rgb", "sam", "mask"])
def test_fused_equals_composed_on_a_large_batch(mode):
    """Size-independent cross-check at a size the CPU oracle would take minutes for: the fused launch and the
    op-by-op torch path (same CUDA encoders) agree on 40k incoherent rays; ragged N; staged == non-staged."""
    opt, params, specs = make_case(with_sam=mode == "sam", with_mask=mode == "mask", max_ray_batch=4096)
    model = build_model(opt, params)
    g = torch.Generator().manual_seed(4)
    N = 40003 if mode == "rgb" else 6007
    pix = torch.randint(0, 800 * 800, (N,), generator=g)
    ro, rd = [], []
    for k in range(4):
        o, d = frame_rays(800, 800, pose_k=k * 5)
        sel = pix[k::4]
        ro.append(o[sel])
        rd.append(d[sel])
    rays_o, rays_d = torch.cat(ro), torch.cat(rd)
    N = rays_o.shape[0]
    kw = {}
    if mode == "sam":
        kw.update(return_feats=1, H=1, W=N)
    if mode == "mask":
        kw.update(return_mask=1)
    a = _render(model, rays_o, rays_d, False, True, **kw)
    b = _render(model, rays_o, rays_d, False, False, **kw)
    assert set(a) == set(b)
    for k in a:
        # both sides are fp32 with different summation orders; near-tie index flips move one sample slightly
        assert_close(a[k], b[k], REL_TOL, f"{mode}/{k}")
    if mode != "sam":
        c = _render(model, rays_o, rays_d, True, True, **kw)
        for k in a:
            assert torch.equal(a[k], c[k]), k          # chunking must not change per-ray arithmetic
    # render quirk: `render(staged=False)` swallows cam_near_far and does NOT forward it to run() (reference renderer.py:187-188),
    # while the staged loop honours it (renderer.py:197-205).  Same grad mode / path for all three renders.
    if mode != "sam":
        cnf = torch.tensor([[0.5, 2.0]], device=DEV)
        ro64, rd64 = rays_o[:64].to(DEV), rays_d[:64].to(DEV)
        model.fused = True
        with torch.no_grad():
            plain = model.render(ro64, rd64, staged=False, **kw)
            dropped = model.render(ro64, rd64, staged=False, cam_near_far=cnf, **kw)
            honoured = model.render(ro64, rd64, staged=True, cam_near_far=cnf, **kw)
        for k in plain:
            assert torch.equal(plain[k], dropped[k]), k
        assert not torch.equal(honoured["depth"], plain["depth"])
        assert float(honoured["depth"].max()) <= 2.0 * (1 + 1e-5)


def test_object_head_chunking_is_invisible():
    """`_run_fused` renders frames larger than its scratch in chunks of rays (one render + one head launch each, pointers
    advanced per chunk): a small `mask_chunk_rays` must reproduce the single-chunk result bit for bit, ragged last chunk
    included, with per-ray near/far and background rows."""
    opt, params, specs = make_case(with_mask=True)
    model = build_model(opt, params)
    rays_o, rays_d = frame_rays(800, 800, pose_k=3)
    g = torch.Generator().manual_seed(11)
    sel = torch.randint(0, 800 * 800, (5003,), generator=g)
    rays_o, rays_d = rays_o[sel].to(DEV), rays_d[sel].to(DEV)
    cnf = torch.stack([torch.full((5003,), 0.3), torch.rand(5003, generator=g) * 3 + 2], dim=-1).to(DEV)
    bg = torch.rand(5003, 3, generator=g).to(DEV)
    model.eval()
    model.fused = True
    with torch.no_grad():
        one = model.run(rays_o, rays_d, bg_color=bg, cam_near_far=cnf, return_mask=1)
        model.opt.mask_chunk_rays = 1024
        many = model.run(rays_o, rays_d, bg_color=bg, cam_near_far=cnf, return_mask=1)
        model.opt.mask_chunk_rays = 0
    assert set(one) == set(many) and "instance_mask_logits" in one
    for k in one:
        assert torch.equal(one[k], many[k]), k


def test_empty_and_tiny_batches():
    opt, params, specs = make_case(small=True)
    model = build_model(opt, params, small=True)
    rays_o, rays_d = frame_rays(32, 32)
    for n in (0, 1, 15, 17):
        out = _render(model, rays_o[:n], rays_d[:n], True, True)
        assert out["image"].shape == (n, 3) and out["depth"].shape == (n,)
    ref, _ = O.run(params, specs, opt, rays_o[:17], rays_d[:17])
    out = _render(model, rays_o[:17], rays_d[:17], True, True)
    assert_close(out["image"], ref["image"], REL_TOL, "image")
    # rays that miss the AABB (origin far outside, pointing away): near=far=1e9 -> reference semantics (inf/NaN scrubbed)
    ro = torch.tensor([[500.0, 0, 0]]).repeat(8, 1)
    rd = torch.tensor([[1.0, 0.1, 0.2]]).repeat(8, 1)
    ref, _ = O.run(params, specs, opt, ro, rd)
    out = _render(model, ro, rd, True, True)
    assert torch.isfinite(out["image"]).all() == torch.isfinite(ref["image"]).all()
    good = torch.isfinite(ref["image"])
    assert_close(out["image"].cpu()[good], ref["image"][good], REL_TOL, "miss/image")


def test_training_mode_composed_path_backward():
    """rgb training step shape: composed path returns the extra keys and gradients reach every parameter group."""
    opt, params, specs = make_case(small=True)
    model = build_model(opt, params, small=True).train()
    rays_o, rays_d = frame_rays(32, 32)
    torch.manual_seed(0)
    out = model.render(rays_o[:256].to(DEV), rays_d[:256].to(DEV), staged=False, perturb=True, update_proposal=True)
    for k in ("image", "depth", "weights_sum", "num_points", "weights", "proposal_loss", "distort_loss"):
        assert k in out, k
    assert out["num_points"] == 256 * 32
    loss = out["image"].pow(2).mean() + out["proposal_loss"] + 0.02 * out["distort_loss"]
    loss.backward()
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert model.grid.embeddings.grad.abs().sum() > 0 and model.prop_encoders[0].embeddings.grad.abs().sum() > 0


@pytest.mark.parametrize("mode", ["rgb", "mask"])
def test_render_image_generates_the_rays_in_kernel(mode):
    """`render_image(pose, intrinsics, H, W)` derives the rays inside the fused kernel (reference
    nerf/utils.py::get_rays, full-image branch) and can emit the 8-bit image of trainer.py:1140-1143.  Must agree with
    rendering explicitly generated rays: the ray arithmetic is the same up to the matmul's summation order (<= 1 ulp on the
    directions, which the ill-conditioned far-field spacing amplifies to ~2e-4 on the logits), so outputs agree well inside
    the 1e-3 path tolerance; the uint8 image is the exact cast of the float image."""
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    opt, params, specs = make_case(with_mask=mode == "mask")
    model = build_model(opt, params)
    H, W = 40, 64
    pose, intr = orbit_pose(7), lego_intrinsics(H, W)
    kw = dict(return_mask=1) if mode == "mask" else {}
    for rows in (None, (8, 24)):
        r0, r1 = rows or (0, H)
        rays_o, rays_d = get_rays(pose, intr, H, W, rows=(r0, r1))
        want = _render(model, rays_o, rays_d, True, True, **kw)
        got = model.render_image(pose, intr, H, W, rows=rows, return_uint8=True, **kw)
        assert set(want) | {"image_u8"} == set(got)
        for k in want:
            assert got[k].shape == want[k].shape
            assert_close(got[k], want[k], 5e-4, f"{mode}/{k}")
        assert got["image_u8"].dtype == torch.uint8 and tuple(got["image_u8"].shape) == ((r1 - r0) * W, 3)
        assert torch.equal(got["image_u8"], (got["image"] * 255).clamp(0, 255).to(torch.uint8))
    # in-kernel rays == reference get_rays formula evaluated in float64 (pixel centres, flipped y/z, unnormalised)
    rays_o, rays_d = get_rays(pose, intr, H, W)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float64) + 0.5, torch.arange(W, dtype=torch.float64) + 0.5, indexing="ij")
    fx, fy, cx, cy = intr
    dirs = torch.stack(((i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)), -1).reshape(-1, 3)
    want_d = dirs @ pose[:3, :3].double().t()
    assert (rays_d.double() - want_d).abs().max() < 1e-6
    # perturbed sampling through the same entry point: seeded, and different from the unperturbed frame
    torch.manual_seed(1)
    p1 = model.render_image(pose, intr, H, W, perturb=True, **kw)
    torch.manual_seed(1)
    p2 = model.render_image(pose, intr, H, W, perturb=True, **kw)
    assert torch.equal(p1["image"], p2["image"]) and not torch.equal(p1["depth"], model.render_image(pose, intr, H, W, **kw)["depth"])
    if mode == "rgb":
        model.train()
        with pytest.raises(RuntimeError):      # rgb training mode also returns losses: not a render_image case
            model.render_image(pose, intr, H, W)
        model.eval()


def test_per_ray_near_far_and_background_match_oracle():
    """cam_near_far [N,2] and bg_color [N,3] per ray (renderer.py:197-205, 233-235, 353), through the staged entry point."""
    opt, params, specs = make_case()
    model = build_model(opt, params)
    rays_o, rays_d = frame_rays(800, 800, pose_k=9, rows=(400, 408), cols=(100, 164))   # 512 rays
    N = rays_o.shape[0]
    g = torch.Generator().manual_seed(2)
    cnf = torch.stack([0.2 + torch.rand(N, generator=g), 3.0 + 4.0 * torch.rand(N, generator=g)], dim=-1)
    bg = torch.rand(N, 3, generator=g)
    ref, _ = O.run(params, specs, opt, rays_o, rays_d, bg_color=bg, cam_near_far=cnf)
    out = _render(model, rays_o, rays_d, True, True, bg_color=bg.to(DEV), cam_near_far=cnf.to(DEV))
    for k, v in ref.items():
        assert_close(out[k], v, REL_TOL, f"per-ray/{k}")


def test_full_frame_properties_and_determinism():
    """BASELINE config #2 at full size (800x800, 640 000 rays, one launch): size-independent properties -- bit-identical
    repeat (no atomics on the path), any split of the ray list gives the same pixels (rays are independent: this is what
    makes the multi-GPU sharding exact), weights sum to 1 with the opaque last sample, outputs finite, depth inside
    [near, far]."""
    opt, params, specs = make_case()
    model = build_model(opt, params)
    rays_o, rays_d = frame_rays(800, 800, pose_k=1)
    a = _render(model, rays_o, rays_d, True, True)
    b = _render(model, rays_o, rays_d, True, True)
    for k in a:
        assert torch.equal(a[k], b[k]), k
        assert torch.isfinite(a[k]).all(), k
    # ragged 3-way split (what parallel.shard_bounds would hand to 3 ranks) == the whole frame, bit for bit
    N = rays_o.shape[0]
    cuts = [0, N // 3 + 1, 2 * N // 3 + 5, N]
    parts = [_render(model, rays_o[cuts[i]:cuts[i + 1]], rays_d[cuts[i]:cuts[i + 1]], True, True) for i in range(3)]
    for k in a:
        assert torch.equal(torch.cat([p[k] for p in parts]), a[k]), k
    # the tile-traversal hint only changes the order in which rays are processed
    c = _render(model, rays_o, rays_d, True, True, image_width=800)
    for k in a:
        assert torch.equal(a[k], c[k]), k
    assert float((a["weights_sum"] - 1).abs().max()) < 1e-5          # background == 'last_sample': the last sample is opaque
    assert float(a["depth"].min()) >= opt.min_near * 0.999
    assert float(a["image"].min()) >= 0.0 and float(a["image"].max()) <= 1.0 + 1e-5


@pytest.mark.parametrize("mode", ["rgb", "mask"])
def test_fused_perturb_consumes_the_reference_random_stream(mode):
    """perturb=True (renderer.py:267-270, 99-100; trainer.py:513 renders the SAM stage's RGB frames this way, the GUI its spp
    frames) in the fused kernel: the jitter is drawn with torch.rand in the reference's order and chunking, so under the same
    seed the fused launch and the op-by-op path (whose agreement with the reference under a shared seed
    tests/test_trainer_boundary_gpu.py checks) see the same random numbers and must agree like the unperturbed renders do."""
    opt, params, specs = make_case(with_mask=mode == "mask", max_ray_batch=2048)
    model = build_model(opt, params)
    g = torch.Generator().manual_seed(8)
    rays_o, rays_d = frame_rays(800, 800, pose_k=6)
    sel = torch.randint(0, 800 * 800, (5003,), generator=g)
    rays_o, rays_d = rays_o[sel].contiguous(), rays_d[sel].contiguous()
    kw = dict(return_mask=1) if mode == "mask" else {}
    for staged in (True, False):            # staged: three chunks of the random stream; non-staged: one
        torch.manual_seed(3)
        a = _render(model, rays_o, rays_d, staged, True, perturb=True, **kw)
        torch.manual_seed(3)
        b = _render(model, rays_o, rays_d, staged, False, perturb=True, **kw)
        torch.manual_seed(3)
        a2 = _render(model, rays_o, rays_d, staged, True, perturb=True, **kw)
        torch.manual_seed(4)
        c = _render(model, rays_o, rays_d, staged, True, perturb=True, **kw)
        plain = _render(model, rays_o, rays_d, staged, True, **kw)
        assert set(a) == set(b)
        for k in a:
            assert_close(a[k], b[k], REL_TOL, f"perturb/{mode}/{k}")
            assert torch.equal(a[k], a2[k]), k
        assert not torch.equal(a["depth"], c["depth"]) and not torch.equal(a["depth"], plain["depth"])


@pytest.mark.parametrize("mode", ["sam", "mask"])
def test_head_training_on_frozen_geometry_runs_the_fused_kernel(mode):
    """Object / SAM stage training after main.py's name-based freezing (main.py:249-256; trainer.py:401-409, 507-550): the
    geometry comes from the fused kernel (no-grad), only the head's feature grid + MLP run as differentiable ops on the samples
    it placed.  Outputs and gradients must match the op-by-op path."""
    from sanerf_hq_b200 import _lib
    opt, params, specs = make_case(with_sam=mode == "sam", with_mask=mode == "mask")
    model = build_model(opt, params).train()
    heads = ("s_grid", "samvit_mlp") if mode == "sam" else ("m_grid", "mask_mlp")
    for n, q in model.named_parameters():
        q.requires_grad = n.startswith(heads)
    g = torch.Generator().manual_seed(12)
    rays_o, rays_d = frame_rays(800, 800, pose_k=10)
    sel = torch.randint(0, 800 * 800, (1024,), generator=g)
    rays_o, rays_d = rays_o[sel].contiguous().to(DEV), rays_d[sel].contiguous().to(DEV)
    kw = dict(return_feats=1, H=32, W=32) if mode == "sam" else dict(return_mask=1)
    key = "samvit" if mode == "sam" else "instance_mask_logits"
    target = torch.randn(1024, 256 if mode == "sam" else 2, generator=g).to(DEV)

    def step(fused):
        model.fused = fused
        model.zero_grad(set_to_none=True)
        n0 = _lib.launch_counter["n"]
        out = model.render(rays_o, rays_d, staged=False, perturb=False, update_proposal=False, **kw)
        used_fused_kernel = "render" if fused else None
        loss = (out[key].reshape(target.shape) - target).pow(2).mean()
        loss.backward()
        grads = {n: q.grad.detach().clone() for n, q in model.named_parameters() if q.grad is not None}
        return out, float(loss), grads, _lib.launch_counter["n"] - n0, used_fused_kernel

    a, la, ga, launches, _ = step(True)
    b, lb, gb, _, _ = step(False)
    model.fused = True
    assert launches >= 2 and all(n.startswith(heads) for n in ga) and set(ga) == set(gb)
    assert set(a) == set(b)
    for k in a:
        assert_close(a[k], b[k], REL_TOL, f"frozen/{mode}/{k}")
    assert abs(la - lb) <= 1e-4 * abs(lb)
    # gradients: the two paths place the samples / evaluate geo_feat with different (split-precision vs cuBLAS) arithmetic, and
    # leaky_relu's derivative is discontinuous where a pre-activation is ~0, so single entries may move by a percent of the scale
    for n in ga:
        scale = float(gb[n].abs().max())
        assert scale > 0 and float((ga[n] - gb[n]).abs().max()) <= 2e-2 * scale, n
        assert float((ga[n] - gb[n]).norm()) <= 1e-2 * float(gb[n].norm()), n
    # with a trainable geometry parameter the call must fall back to the fully differentiable path
    model.grid_mlp.net[0].weight.requires_grad = True
    assert not model._can_train_heads_on_fused_geometry(rays_o, dict(kw, perturb=False))


def test_exact_early_out_behind_opaque_samples():
    """The proposal stage skips positions / gathers / density of sample chunks whose transmittance is exactly 0 in fp32 (the
    delta*sigma in front sums to > 105): their weights are exactly 0 either way.  A proposal network with a steep density makes
    most rays opaque within the first 64 coarse samples; every output and both sample_pdf index buffers must still match the
    oracle, which evaluates everything."""
    opt, params, specs = make_case()
    params = {k: v.clone() for k, v in params.items()}
    params["prop_mlp.0.net.1.weight"] *= 100.0        # ~half of the rays become opaque within the first 64 coarse samples
    model = build_model(opt, params)
    rays_o, rays_d = frame_rays(800, 800, pose_k=8, rows=(380, 396), cols=(300, 332))   # 512 rays
    ref, ex = O.run(params, specs, opt, rays_o, rays_d)
    ds0 = (ex["real_bins"][0][:, 1:] - ex["real_bins"][0][:, :-1]) * ex["sigmas"][0]
    opaque_early = (ds0[:, :64].sum(-1) > 105).float().mean().item()
    assert opaque_early > 0.2, f"the test scene must exercise the early-out ({opaque_early:.2f} of the rays do)"
    taps = dict(inds0=None, inds1=None, weights2=None, bins2=None)
    out = _render(model, rays_o, rays_d, False, True, taps=taps)
    for k, v in ref.items():
        assert_close(out[k], v, REL_TOL, f"early-out/{k}")
    assert_close(taps["bins2"], ex["bins"][2], REL_TOL, "early-out/bins2")
    for i, name in enumerate(("inds0", "inds1")):
        aux = ex["pdf"][i]
        n_bad, n_unexpl = index_mismatch_report(taps[name].cpu().numpy(), aux["inds"].numpy(), aux["cdf"].numpy(), aux["u"].numpy())
        assert n_unexpl == 0, (name, n_bad, n_unexpl)
