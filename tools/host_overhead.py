#!/usr/bin/env python
"""Host-side cost of one fused render call (GPU box): wall time per `model.render` of a small ray batch, where the GPU work
(a few tens of microseconds) hides nothing.  Prints microseconds per call for rgb / mask / sam."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose  # noqa: E402

dev = "cuda"
ro, rd = get_rays(orbit_pose(1).to(dev), lego_intrinsics(16, 16), 16, 16, device=dev)
with torch.no_grad():
    for wl in ("rgb", "mask", "sam"):
        model = bench.build_model(wl, dev)
        kw = dict(return_mask=1) if wl == "mask" else (dict(return_feats=1, H=16, W=16) if wl == "sam" else {})
        for _ in range(20):
            model.render(ro, rd, staged=wl != "sam", perturb=False, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 300
        for _ in range(n):
            model.render(ro, rd, staged=wl != "sam", perturb=False, **kw)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{wl}: {1e6 * (t1 - t0) / n:.0f} us per call on the host (256 rays; + {1e6 * (t2 - t1):.0f} us to drain the queue)")
        del model
