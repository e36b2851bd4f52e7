#!/usr/bin/env python
"""Diagnose the worst ray of a full-frame comparison against the reference GPU render (run on the GPU box):
which of {candidate, reference GPU, CPU oracle} disagree on it, and where along the pipeline (taps).

    python tools/diag_outlier.py [sam|mask] [pose]
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import bench  # noqa: E402
from oracle import ref_runtime as R  # noqa: E402
from oracle import render_oracle as O  # noqa: E402
from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sam"
pose = int(sys.argv[2]) if len(sys.argv) > 2 else 11
key = "samvit" if wl == "sam" else "instance_mask_logits"
DEV, H, W = "cuda", 800, 800
torch.backends.cuda.matmul.allow_tf32 = False
cand = bench.build_model(wl, DEV)
ref = R.build_network(bench.default_opt(wl), cand.state_dict(), device=DEV)
ref.opt.max_ray_batch = 16384
ro, rd = get_rays(orbit_pose(pose).to(DEV), lego_intrinsics(H, W), H, W, device=DEV)
with torch.no_grad():
    if wl == "sam":
        want = R.render_features_by_rows(ref, ro, rd, W, rows_per_call=5, perturb=False, bg_color=1)
        got = cand.render(ro, rd, staged=False, perturb=False, bg_color=1, return_feats=1, H=H, W=W, image_width=W)
    else:
        want = R.render(ref, ro, rd, staged=True, perturb=False, bg_color=1, return_mask=1)
        got = cand.render(ro, rd, staged=True, perturb=False, bg_color=1, return_mask=1)
a, b = got[key].reshape(H * W, -1).double(), want[key].reshape(H * W, -1).double()
floor = max(1e-3, 0.1 * float(b.pow(2).mean().sqrt()))
e = ((a - b).abs() / b.abs().clamp(min=floor)).amax(dim=1)
worst = torch.topk(e, 8)
print("floor", floor, "worst rays", worst.indices.tolist(), [f"{v:.2e}" for v in worst.values.tolist()])
specs = O.default_specs(2)
params = {k: v.detach().cpu() for k, v in cand.state_dict().items()}
opt = bench.default_opt(wl)
for r in worst.indices.tolist()[:3]:
    sel = torch.tensor([r, r, r, r], device=DEV)       # 4 copies: whole 128-sample tile
    o, d = ro[sel].contiguous(), rd[sel].contiguous()
    kw = dict(return_feats=1, H=1, W=4) if wl == "sam" else dict(return_mask=1)
    with torch.no_grad():
        g = R.render(ref, o, d, staged=False, perturb=False, bg_color=1, **kw)
        taps = dict(inds0=None, inds1=None, weights2=None, sigma2=None, bins2=None, f_image=None)
        c = cand._run_fused(o, d, bg_color=1, taps=taps, **kw)
    cpu, ex = O.run(params, specs, opt, o.cpu(), d.cpu(), bg_color=1, **kw)
    print(f"--- ray {r} (row {r // W}, col {r % W})  frame err {float(e[r]):.3e}")
    for k in ("image", "depth", "weights_sum", key):
        x = {"cand": c[k].reshape(4, -1)[0].double().cpu(), "refgpu": g[k].reshape(4, -1)[0].double().cpu(), "cpu": cpu[k].reshape(4, -1)[0].double()}
        fl = floor if k == key else 1e-3
        def err(p, q):
            return float(((x[p] - x[q]).abs() / x[q].abs().clamp(min=fl)).max())
        print(f"  {k:22s} cand-refgpu {err('cand', 'refgpu'):.2e}  cand-cpu {err('cand', 'cpu'):.2e}  refgpu-cpu {err('refgpu', 'cpu'):.2e}")
    i0, i1 = taps["inds0"][0].cpu().long(), taps["inds1"][0].cpu().long()
    print("  inds0 mismatches vs cpu", int((i0 != ex["pdf"][0]["inds"][0]).sum()), " inds1", int((i1 != ex["pdf"][1]["inds"][0]).sum()))
    print("  bins2 max abs diff", float((taps["bins2"][0].cpu() - ex["bins"][2][0]).abs().max()),
          " weights2 max abs diff", float((taps["weights2"][0].cpu() - ex["weights"][2][0]).abs().max()))
    w2, wc = taps["weights2"][0].cpu(), ex["weights"][2][0]
    print("  weights2 cand", [f"{v:.3e}" for v in w2.tolist()])
    print("  weights2 cpu ", [f"{v:.3e}" for v in wc.tolist()])
    s2 = taps["sigma2"][0].cpu()
    print("  sigma2 cand", [f"{v:.3e}" for v in s2.tolist()])
    if "sigmas" in ex:
        print("  sigma2 cpu ", [f"{v:.3e}" for v in ex["sigmas"][2][0].tolist()])
