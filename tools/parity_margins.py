import sys, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import *
import numpy as np, os
for name in ["cfg1_rgb", "cfg1_sam", "cfg1_mask", "full_rgb", "full_sam", "full_mask"]:
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    small, with_sam, with_mask, h, w, batch, staged = [int(v) for v in fx["meta"]]
    opt, params, specs = make_case(small=bool(small), with_sam=bool(with_sam), with_mask=bool(with_mask), max_ray_batch=batch)
    model = build_model(opt, params, small=bool(small))
    kw = dict(perturb=False, bg_color=1)
    if with_sam: kw.update(return_feats=1, H=h, W=w)
    if with_mask: kw.update(return_mask=1)
    with torch.no_grad():
        out = model.render(torch.from_numpy(fx["rays_o"]).cuda(), torch.from_numpy(fx["rays_d"]).cuda(), staged=bool(staged), **kw)
    print(name, {k: f"{rel_err(out[k], fx['out_' + k], floor=tol_floor(name + '/' + k, fx['out_' + k])):.2e}" for k in out if torch.is_tensor(out[k])})
for mode in ["rgb", "sam", "mask"]:
    opt, params, specs = make_case(with_sam=mode == "sam", with_mask=mode == "mask", table_scale=3.0)
    model = build_model(opt, params)
    rays_o, rays_d = frame_rays(800, 800, pose_k=5, rows=(200, 216), cols=(384, 416))
    kw = {}
    if mode == "sam": kw.update(return_feats=1, H=16, W=32)
    if mode == "mask": kw.update(return_mask=1)
    ref, ex = O.run(params, specs, opt, rays_o, rays_d, **kw)
    taps = dict(weights2=None, sigma2=None)
    with torch.no_grad():
        out = model._run_fused(rays_o.cuda(), rays_d.cuda(), taps=taps, **kw)
    print(mode, "scale3", {k: f"{rel_err(out[k], v, floor=tol_floor(mode + '/' + k, v)):.2e}" for k, v in ref.items()},
          "weights2", f"{rel_err(taps['weights2'], ex['weights'][2]):.2e}", "sigma2", f"{rel_err(taps['sigma2'], ex['sigmas'][2] if 'sigmas' in ex else taps['sigma2']):.2e}")
