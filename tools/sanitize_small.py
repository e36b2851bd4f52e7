#!/usr/bin/env python
"""Small renders of every fused variant for compute-sanitizer (run under `compute-sanitizer --tool memcheck|racecheck`):
default-size networks, 96 rays, rgb / rgb perturbed / rgb with early-out rays / SAM feature (NHWC, NCHW + resize) / object head / frozen-geometry training."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose  # noqa: E402

dev = "cuda"
ro, rd = get_rays(orbit_pose(2).to(dev), lego_intrinsics(8, 12), 8, 12, device=dev)
for wl in ("rgb", "sam", "mask"):
    model = bench.build_model(wl, dev)
    with torch.no_grad():
        if wl == "rgb":
            model.render(ro, rd, staged=True, perturb=False)
            model.render(ro, rd, staged=True, perturb=True)
            model.render_image(orbit_pose(1), lego_intrinsics(8, 12), 8, 12, return_uint8=True)
            # a steep-density scene: most rays turn opaque inside the first 64 coarse samples -> the exact early-out of the proposal stage
            model.prop_mlp[0].net[1].weight.mul_(100.0)
            model.render(ro, rd, staged=True, perturb=False)
            model.prop_mlp[0].net[1].weight.div_(100.0)
        elif wl == "sam":
            model.render(ro, rd, staged=False, perturb=False, return_feats=1, H=8, W=12)
            model.render(ro, rd, staged=False, perturb=False, return_feats=1, H=8, W=12, feature_layout="nchw", feature_size=(5, 7))
        else:
            model.render(ro, rd, staged=True, perturb=False, return_mask=1)
            model.render(ro, rd, staged=True, perturb=True, return_mask=1)
    if wl != "rgb":
        model.train()
        heads = ("s_grid", "samvit_mlp") if wl == "sam" else ("m_grid", "mask_mlp")
        for n, q in model.named_parameters():
            q.requires_grad = n.startswith(heads)
        kw = dict(return_feats=1, H=8, W=12) if wl == "sam" else dict(return_mask=1)
        out = model.render(ro, rd, staged=False, perturb=False, update_proposal=False, **kw)
        out["samvit" if wl == "sam" else "instance_mask_logits"].pow(2).mean().backward()
    torch.cuda.synchronize()
    print(wl, "ok")
    del model
