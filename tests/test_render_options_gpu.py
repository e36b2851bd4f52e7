"""GPU: the fused and the composed render against the option-variant fixtures generated from the reference
(tests/golden/make_golden.py: --background white, contract=False / bound=1, per-ray cam_near_far through the staged loop) and
against the oracle on adversarial rays.  The oracle is pinned on the same fixtures by tests/test_oracle_golden.py (CPU)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, REL_TOL, assert_close, build_model, make_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "composed"])
@pytest.mark.parametrize("name", ["opt_white", "opt_box", "opt_cnf"])
def test_render_matches_reference_option_fixture(name, fused):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    small, with_sam, with_mask, h, w, batch, staged = [int(v) for v in fx["meta"]]
    optkw = json.loads(str(fx["optkw"])) if "optkw" in fx.files else {}
    opt, params, specs = make_case(small=bool(small), with_sam=bool(with_sam), with_mask=bool(with_mask), max_ray_batch=batch, **optkw)
    model = build_model(opt, params, small=bool(small))
    model.fused = fused
    rays_o, rays_d = torch.from_numpy(fx["rays_o"]).to(DEV), torch.from_numpy(fx["rays_d"]).to(DEV)
    kw = dict(perturb=False, bg_color=1)
    if "cam_near_far" in fx.files:
        kw["cam_near_far"] = torch.from_numpy(fx["cam_near_far"]).to(DEV)
    with torch.no_grad():
        out = model.render(rays_o, rays_d, staged=bool(staged), **kw)
    want_keys = {k[4:] for k in fx.files if k.startswith("out_")}
    assert {k for k, v in out.items() if torch.is_tensor(v)} == want_keys
    for k in sorted(want_keys):
        assert out[k].shape == fx["out_" + k].shape, k
        assert_close(out[k], fx["out_" + k], REL_TOL, f"{name}/{k}")


def test_fused_handles_adversarial_rays_like_the_oracle():
    """Axis-parallel directions (the +1e-15 guard of near_far_from_aabb), origins inside the box, rays that miss it: the rays of
    tests/golden/helpers.npz (whose near/far the reference itself produced) through the fused kernel vs the oracle."""
    from helpers import O
    fx = np.load(os.path.join(GOLDEN, "helpers.npz"))
    rays_o, rays_d = torch.from_numpy(fx["rays_o"]), torch.from_numpy(fx["rays_d"])
    for optkw in ({}, dict(contract=False, bound=1)):
        opt, params, specs = make_case(small=True, **optkw)
        model = build_model(opt, params, small=True)
        model.fused = True
        ref, _ = O.run(params, specs, opt, rays_o, rays_d)
        with torch.no_grad():
            out = model.run(rays_o.to(DEV), rays_d.to(DEV))
        for k, v in ref.items():
            got, nan = out[k].cpu(), torch.isnan(v)       # rays that miss a bound=1 box: the reference's depth is 0 * 1e9-ish = NaN
            assert torch.equal(torch.isnan(got), nan), f"{optkw}/{k}: NaN pattern differs from the reference's"
            assert_close(got[~nan], v[~nan], REL_TOL, f"{optkw}/{k}")
