"""Shared test utilities: build a NeRFNetwork with the oracle's seeded weights, tolerances, tie-aware
index comparison."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import render_oracle as O  # noqa: E402

# SURVEY.md 8d tolerance: |cand - ref| <= 1e-3 * max(|ref|, 1e-3)
REL_TOL = 1e-3


def rel_err(cand, ref, floor=1e-3):
    cand = torch.as_tensor(cand).detach().double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).detach().double().cpu().reshape(-1)
    return ((cand - ref).abs() / ref.abs().clamp(min=floor)).max().item()


# Signed dot-product outputs (composited feature vectors, LayerNorm'd SAM features, logits) cancel to ~0 in some
# channels; an fp32 dot product's error scales with sum|w_i x_i|, not with |sum w_i x_i|, so for those keys the
# floor of the relative error is 10 % of the tensor's rms instead of the 1e-3 used for image / depth / weights:
#     |cand - ref| <= 1e-3 * max(|ref|, 0.1 * rms(ref))
# The internal tap f_image (never returned to the caller; its consumer `image` is checked at the strict floor) is
# compared against its rms: a +-1 ulp change of the resampled bins (parallel-scan vs serial cumsum in sample_pdf)
# already moves it by 3.5e-4 of 0.1*rms in the oracle itself (ill-conditioned 1/(2-2x) spacing at far ~ 200).
SIGNED_VECTOR_KEYS = ("samvit", "f_image", "instance_mask_logits")
RMS_FRACTION = {"f_image": 1.0}


def tol_floor(what, ref):
    if what.split("/")[-1] in SIGNED_VECTOR_KEYS:
        r = torch.as_tensor(ref).detach().double()
        return max(1e-3, RMS_FRACTION.get(what.split("/")[-1], 0.1) * float(r.pow(2).mean().sqrt()))
    return 1e-3


def assert_close(cand, ref, tol=REL_TOL, what=""):
    e = rel_err(cand, ref, floor=tol_floor(what, ref))
    assert e <= tol, f"{what}: max rel err {e:.3e} > {tol:.1e}"
    return e


def make_case(small=False, with_sam=False, with_mask=False, seed=7, table_scale=1.0, **optkw):
    """(opt, params, specs) exactly as tests/golden/make_golden.py builds them."""
    opt = O.default_opt(with_sam=with_sam, with_mask=with_mask, **optkw)
    specs = O.default_specs(2 if opt.contract else opt.bound, num_levels=4 if small else None)
    params, specs = O.make_params(opt, specs, seed=seed, hidden=16 if small else None, table_scale=table_scale)
    return opt, params, specs


def build_model(opt, params, small=False, device="cuda"):
    """This repo's NeRFNetwork loaded with the oracle's weights (reference state_dict key names)."""
    from sanerf_hq_b200.network import NeRFNetwork
    kw = dict(num_levels=4, hidden_dim=16) if small else {}
    model = NeRFNetwork(opt, **kw)
    res = model.load_state_dict(params, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.eval().to(device)


def frame_rays(H, W, pose_k=3, rows=None, cols=None):
    rays_o, rays_d = O.get_rays(O.orbit_pose(pose_k), H, W)
    if rows is not None:
        idx = (torch.arange(rows[0], rows[1])[:, None] * W + torch.arange(cols[0], cols[1])[None, :]).reshape(-1)
        rays_o, rays_d = rays_o[idx].contiguous(), rays_d[idx].contiguous()
    return rays_o, rays_d


def param_digest(params):
    return float(sum(float(v.double().abs().sum()) for k, v in sorted(params.items()) if v.is_floating_point()))


def index_mismatch_report(cand, ref, cdf, u, ulps=4):
    """Compare searchsorted index buffers.  A mismatch is 'explained' when the oracle's cdf has an
    entry within `ulps` fp32 ulps (plus a 1e-6 band for upstream fp32 summation-order noise in the
    weights) of the query u -- the only situation in which two correct fp32 evaluations of
    searchsorted(cdf, u) can disagree.  Returns (n_mismatch, n_unexplained)."""
    cand = np.asarray(cand).astype(np.int64)
    ref = np.asarray(ref).astype(np.int64)
    bad = np.argwhere(cand != ref)
    unexplained = 0
    for r, k in bad:
        if abs(int(cand[r, k]) - int(ref[r, k])) > 1:
            unexplained += 1
            continue
        j = max(cand[r, k], ref[r, k]) - 1          # the cdf entry the two sides disagree about
        c, uu = float(cdf[r, j]), float(u[r, k]) if np.ndim(u) == 2 else float(u[k])
        tol = ulps * np.spacing(np.float32(max(abs(c), abs(uu)))) + 2e-6
        if abs(c - uu) > tol:
            unexplained += 1
    return len(bad), unexplained
