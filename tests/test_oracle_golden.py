"""CPU: the oracle (oracle/) reproduces the golden fixtures generated from the reference itself
(tests/golden/make_golden.py).  This pins the oracle; the GPU tests then compare CUDA vs oracle."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, make_case, param_digest, rel_err

CASES = ["cfg1_rgb", "cfg1_sam", "cfg1_mask", "full_rgb", "full_sam", "full_mask",
         "opt_white", "opt_box", "opt_cnf"]      # option variants: --background white, contract=False / bound=1, per-ray near/far


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    small, with_sam, with_mask, h, w, batch, staged = [int(v) for v in fx["meta"]]
    optkw = json.loads(str(fx["optkw"])) if "optkw" in fx.files else {}
    cnf = torch.from_numpy(fx["cam_near_far"]) if "cam_near_far" in fx.files else None
    opt, params, specs = make_case(small=bool(small), with_sam=bool(with_sam), with_mask=bool(with_mask), max_ray_batch=batch, **optkw)
    # the seeded weights must be the ones the fixture was generated with
    assert abs(param_digest(params) - float(fx["param_digest"])) <= 1e-6 * float(fx["param_digest"])
    rays_o, rays_d = torch.from_numpy(fx["rays_o"]), torch.from_numpy(fx["rays_d"])
    kw = dict(bg_color=1)
    if with_sam:
        kw.update(return_feats=1, H=h, W=w)
    if with_mask:
        kw.update(return_mask=1)
    inds0, inds1, outs = [], [], {}
    N = rays_o.shape[0]
    step = batch if staged else N
    for head in range(0, N, step):
        r, ex = O.run(params, specs, opt, rays_o[head:head + step], rays_d[head:head + step],
                      cam_near_far=None if cnf is None else cnf[head:head + step], **kw)
        inds0.append(ex["pdf"][0]["inds"])
        inds1.append(ex["pdf"][1]["inds"])
        for k, v in r.items():
            outs.setdefault(k, []).append(v.reshape(-1, *v.shape[2:]) if k == "samvit" else v)
    for k in outs:
        got = torch.cat(outs[k], 0).reshape(fx["out_" + k].shape)
        # same torch build -> bit-identical in the container that made the fixture; other CPUs may pick
        # different GEMM kernels, hence a (tight) tolerance rather than equality
        assert rel_err(got, fx["out_" + k]) < 2e-5, k
    i0, i1 = torch.cat(inds0).numpy(), torch.cat(inds1).numpy()
    assert (i0 != fx["inds0"]).mean() < 2e-3 and (i1 != fx["inds1"]).mean() < 2e-3
    assert np.abs(i0 - fx["inds0"]).max() <= 1 and np.abs(i1 - fx["inds1"]).max() <= 1


def test_level_resolution_table_matches_survey():
    """Kernel-side fp32 resolutions (gridencoder.cu:133) differ from the host float64 rule at some levels
    (SURVEY.md appendix B)."""
    from oracle import kernels as K
    sp = O.default_specs(2)
    S = np.log2(sp["grid"].per_level_scale)
    res = [K.level_resolution(l, S, 16) for l in range(16)]
    assert res == [16, 24, 34, 49, 71, 102, 148, 213, 308, 446, 646, 934, 1352, 1956, 2831, 4096]
    S = np.log2(sp["s_grid"].per_level_scale)
    res = [K.level_resolution(l, S, 16) for l in range(16)]
    assert res == [16, 21, 26, 32, 41, 51, 64, 81, 102, 128, 162, 204, 256, 323, 407, 512]
    assert sp["grid"].offsets[-1] == 6299960 and sp["s_grid"].offsets[-1] == 5258512
    assert sp["prop_encoders.0"].offsets[-1] == 383264 and sp["prop_encoders.1"].offsets[-1] == 430080


def test_oracle_grid_kernel_edge_cases():
    """OOB -> zeros; x=0 and x=1 hit the clamped first / last cell; dense->hash transition level."""
    from oracle import kernels as K
    sp = O.grid_spec(16, 2, 16, 19, 4096)
    g = torch.Generator().manual_seed(0)
    emb = torch.rand(int(sp.offsets[-1]), 2, generator=g) * 2 - 1
    x = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [1.5, 0.5, 0.5], [-0.1, 0.2, 0.3], [0.5, 0.5, 0.5]])
    S = np.log2(sp.per_level_scale)
    out = K.grid_encode_forward(x, emb, sp.offsets, 5, 3, 2, 16, 16, S, 16)
    assert out.shape == (16, 5, 2)
    assert torch.all(out[:, 2] == 0) and torch.all(out[:, 3] == 0)
    # x=0: pos=clamp(-0.5)=0 -> exactly vertex (0,0,0) of each level = row `offset` (dense) or hash(0,0,0)=0
    for l in range(16):
        assert torch.equal(out[l, 0], emb[int(sp.offsets[l])])
    # x=1 on level 0 (res 16, dense): vertex (15,15,15) -> row 15+15*16+15*256
    assert torch.equal(out[0, 1], emb[15 + 15 * 16 + 15 * 256])
    # partial levels: max_level < L leaves the rest untouched (zero-filled by the caller)
    out2 = K.grid_encode_forward(x, emb, sp.offsets, 5, 3, 2, 16, 3, S, 16)
    assert torch.equal(out2[:3], out[:3]) and torch.all(out2[3:] == 0)


def test_training_losses_match_reference_fixture():
    """renderer.py:17-57, training only: `proposal_loss` against the reference's own function (fixture from make_golden.py) and
    `distort_loss` against the published definition of `torch_efficient_distloss.eff_distloss` (third party, unpinned in
    requirements.txt:21, not installed): sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 delta_i evaluated in float64.  Both helpers
    are device-agnostic torch code, so they are checked here on the CPU; fp32 prefix sums vs float64 brute force: 1e-5."""
    import sanerf_hq_b200.renderer as R
    fx = np.load(os.path.join(GOLDEN, "losses.npz"))
    bins = [torch.from_numpy(fx[f"bins{i}"]) for i in range(3)]
    weights = [torch.from_numpy(fx[f"weights{i}"]) for i in range(3)]
    pl = float(R.proposal_loss(bins, weights))
    dl = float(R.distort_loss(bins[-1], weights[-1]))
    assert abs(pl - float(fx["proposal_loss"])) <= 1e-6 * abs(float(fx["proposal_loss"]))
    assert abs(dl - float(fx["distort_loss"])) <= 1e-5 * abs(float(fx["distort_loss"]))
    # gradients flow to the proposal weights only (the last level is the detached target, renderer.py:50-51)
    ws = [w.clone().requires_grad_(True) for w in weights]
    R.proposal_loss(bins, ws).backward()
    assert ws[0].grad is not None and ws[1].grad is not None and ws[2].grad is None
    assert float(ws[0].grad.abs().sum()) > 0


def test_torch_helpers_match_reference_fixture():
    """`contract` / `uncontract` / `near_far_from_aabb` / `sample_pdf` (renderer.py:60-139) on adversarial inputs, oracle AND
    the drop-in module's own copies against outputs of the reference's functions (tests/golden/helpers.npz)."""
    import sanerf_hq_b200.renderer as R
    fx = np.load(os.path.join(GOLDEN, "helpers.npz"))
    t = lambda k: torch.from_numpy(fx[k])
    same = lambda a, b: torch.allclose(a, b, rtol=1e-6, atol=0, equal_nan=True)
    for mod in (O, R):
        z = mod.contract(t("contract_in"))
        assert same(z, t("contract_out")), mod.__name__
        if hasattr(mod, "uncontract"):
            assert same(mod.uncontract(z), t("uncontract_out")), mod.__name__
        near, far = mod.near_far_from_aabb(t("rays_o"), t("rays_d"), t("aabb"), 0.2)
        assert same(near, t("near")) and same(far, t("far")), mod.__name__
        for T0, T in ((128, 65), (64, 33)):
            out = mod.sample_pdf(t(f"pdf{T0}_bins"), t(f"pdf{T0}_weights"), T)
            out = out[0] if isinstance(out, tuple) else out
            assert same(out, t(f"pdf{T0}_out")), (mod.__name__, T0)


def test_oracle_matches_the_reference_python_run_live():
    """Besides the committed fixtures: when the reference's own Python is staged (oracle/_ref/bytecode, built by
    oracle/stage_ref.py where /root/reference is mounted) it is imported here and run on CPU -- its `_gridencoder` /
    `_shencoder` imports served by the C restatement -- and the oracle's restatement of `run()` must reproduce it exactly,
    for all three workloads, staged and non-staged."""
    import warnings
    from oracle import ref_runtime as R
    if not R.available("cpu"):
        pytest.skip("oracle/_ref/bytecode is not staged (python oracle/stage_ref.py)")
    for with_sam, with_mask in ((False, False), (True, False), (False, True)):
        opt = O.default_opt(with_sam=with_sam, with_mask=with_mask, max_ray_batch=64)
        specs = O.default_specs(2)
        params, specs = O.make_params(opt, specs, seed=11)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = R.build_network(opt, params, device="cpu", backend="cpu")
        rays_o, rays_d = O.get_rays(O.orbit_pose(9), 800, 800)
        sel = torch.arange(0, 150) * 4001 % (800 * 800)
        rays_o, rays_d = rays_o[sel].contiguous(), rays_d[sel].contiguous()
        kw = dict(return_feats=1, H=10, W=15) if with_sam else (dict(return_mask=1) if with_mask else {})
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = R.render(model, rays_o, rays_d, backend="cpu", staged=not with_sam, perturb=False, bg_color=1, **kw)
        got = O.render(params, specs, opt, rays_o, rays_d, staged=not with_sam, bg_color=1, **kw)
        for k, v in want.items():
            if torch.is_tensor(v):
                assert torch.equal(got[k].reshape(v.shape), v), (with_sam, with_mask, k)
