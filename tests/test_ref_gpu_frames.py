"""GPU: this repo's fused render vs the REFERENCE ITSELF running on the same GPU, at the BASELINE shapes.

The GPU oracle of SURVEY.md 8c/8d = the reference's own `NeRFNetwork.render` (nerf/renderer.py:185-385, nerf/network.py,
byte-compiled by oracle/stage_ref.py) on the reference's own CUDA kernels (oracle/_ref/_gridencoder.so, _shencoder.so, compiled
verbatim by oracle/build_ref.py), fp32, TF32 off.  Compared on identical weights and rays:

  * config 2   whole 800x800 RGB frame, reference `render(staged=True)`
  * config 3   whole 800x800 RGB + SAM-feature frame, reference driven in 5-row chunks (staged + return_feats is impossible in
               the reference)
  * config 4   (i) whole 800x800 frame with the object head, (ii) a train-style batch: 6000 random pixels over 24 poses + four
               8x8 local patches (scripts/train_obj_nerf.sh:20,27-30; global rays first, provider.py:982-993)
  * the reference's GPU path against the CPU oracle at the strict and at the rms floor (what fp32 itself can promise)

Tolerance: |cand - ref| <= 1e-3 * max(|ref|, floor), floor = 1e-3 for image / depth / weights_sum; for the signed
feature vectors (samvit, instance_mask_logits) floor = max(1e-3, 0.1 * rms(ref)) -- justified by
`test_reference_gpu_vs_cpu_oracle_floors`, which measures how the reference's own fp32 GPU path scores against the fp32 CPU
oracle at both floors (measured on B200: the reference's GPU samvit is 1.7e-2 away from the CPU oracle at the strict floor, i.e.
the strict floor is not a property of the algorithm in fp32; 2.9e-4 at the rms floor).

Arbitration for the signed feature vectors: the head MLPs amplify a one-ulp difference in a resampled bin (a near-tie in
sample_pdf moves a sample), so
on a whole frame (164 M values) two fp32 evaluations of the reference algorithm disagree beyond 1e-3 on a handful of rays --
measured: on the worst ray of pose 11 the reference's GPU path is 2.6e-2 away from the CPU oracle while the candidate is 2e-4 away
from it; on another ray the reference GPU's own frame and its run on that ray alone differ by 1.3e-3, the candidate agrees with
the latter to 8e-5, and both sit 1.4e-3 from the CPU oracle (tools/diag_outlier.py).  For such rays (at most 2e-5 of the frame)
the candidate must agree within tolerance with at least ONE other fp32 evaluation of the reference algorithm on that ray: the
CPU oracle, or the reference's GPU path run on the ray alone; how far those references are apart on the same rays is recorded
next to it.  The measured margins are written to gpurun_out/ref_gpu_parity.json.
"""
import json
import os

import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu
DEV = "cuda"
H = W = 800
STATS = {}


def _R():
    from oracle import ref_runtime as R
    if not R.available("cuda"):
        pytest.skip("oracle/_ref is not staged (python oracle/stage_ref.py; python oracle/build_ref.py)")
    return R


def _record(name, stats):
    STATS[name] = stats
    out = os.path.join(REPO, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "ref_gpu_parity.json")
        old = {}
        if os.path.exists(path):
            old = json.load(open(path))
        old[name] = stats
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _models(workload):
    """Candidate and reference networks with identical weights: the bench scene (seed-0 constructor, hash tables U(-1,1))."""
    import bench
    R = _R()
    torch.backends.cuda.matmul.allow_tf32 = False
    cand = bench.build_model(workload, DEV)
    ref = R.build_network(bench.default_opt(workload), cand.state_dict(), device=DEV)
    ref.opt.max_ray_batch = 16384          # the reference's own default chunk (main.py:90); per-ray arithmetic does not depend on it
    return R, cand, ref


def _frame(pose_k):
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    return get_rays(orbit_pose(pose_k).to(DEV), lego_intrinsics(H, W), H, W, device=DEV)


@torch.no_grad()
def _errors(cand, ref, floor):
    """max and 99.99th percentile of |cand - ref| / max(|ref|, floor) and the fraction of elements beyond 1e-3, on the GPU."""
    worst, beyond, n = 0.0, 0, 0
    cand, ref = cand.reshape(-1), ref.reshape(-1)
    for head in range(0, ref.numel(), 1 << 24):
        a, b = cand[head:head + (1 << 24)].double(), ref[head:head + (1 << 24)].double()
        e = (a - b).abs() / b.abs().clamp(min=floor)
        e = torch.where(torch.isnan(a) & torch.isnan(b), torch.zeros_like(e), e)   # rays that miss: NaN on both sides
        worst = max(worst, float(e.max()))
        beyond += int((e > 1e-3).sum())
        n += e.numel()
    return {"max": worst, "frac_beyond_1e-3": beyond / max(n, 1), "n": n, "floor": floor}


def _rms_floor(ref):
    return max(1e-3, 0.1 * float(ref.double().pow(2).mean().sqrt()))


def _arbitrate(key, cv, rv, floor, arb):
    """Rays whose `key` deviates from the reference GPU frame beyond 1e-3 (at `floor`): at most 2e-5 of the frame, and each must
    agree within tolerance with another fp32 evaluation of the reference algorithm on that ray -- the CPU oracle, or the
    reference GPU run on the ray alone (its CUB scans / cuBLAS kernels associate differently at another batch shape).
    Returns (number of such rays, worst of the per-ray best agreement, worst error of the other rays, how far the references
    themselves are apart on those rays)."""
    import bench
    from oracle import render_oracle
    wl, cand, ro, rd, kw, R, ref = arb
    n = ro.shape[0]
    a, b = cv.reshape(n, -1).double(), rv.reshape(n, -1).double()
    per_ray = ((a - b).abs() / b.abs().clamp(min=floor)).amax(dim=1)
    bad = torch.nonzero(per_ray > 1e-3).reshape(-1)
    good_max = float(per_ray[per_ray <= 1e-3].max()) if bool((per_ray <= 1e-3).any()) else 0.0
    if bad.numel() == 0:
        return 0, 0.0, good_max, 0.0
    assert bad.numel() <= max(2, int(2e-5 * n)), f"{key}: {bad.numel()} rays deviate from the reference GPU beyond tolerance"
    params = {k: v.detach().cpu() for k, v in cand.state_dict().items()}
    okw = dict(kw)
    if "return_feats" in okw:
        okw.update(H=1, W=int(bad.numel()))
    cpu, _ = render_oracle.run(params, render_oracle.default_specs(2), bench.default_opt(wl), ro[bad].cpu(), rd[bad].cpu(), bg_color=1, **okw)
    c = cpu[key].reshape(bad.numel(), -1).double().to(a.device)
    with torch.no_grad():
        g = R.render(ref, ro[bad].contiguous(), rd[bad].contiguous(), staged=False, perturb=False, bg_color=1, **okw)[key]
    g = g.reshape(bad.numel(), -1).double()

    def err(x, y):
        return ((x - y).abs() / y.abs().clamp(min=floor)).amax(dim=1)

    best = torch.minimum(err(a[bad], c), err(a[bad], g))
    spread = torch.maximum(err(b[bad], c), err(g, c))          # reference GPU (frame / alone) vs CPU oracle on the same rays
    return int(bad.numel()), float(best.max()), good_max, float(spread.max())


def _check(name, cand_out, ref_out, signed=(), arb=None):
    stats = {}
    for k, rv in ref_out.items():
        if not torch.is_tensor(rv):
            continue
        cv = cand_out[k].reshape(rv.shape)
        stats[k] = _errors(cv, rv, 1e-3)
        if k in signed:
            floor = _rms_floor(rv)
            stats[k + "@rms_floor"] = _errors(cv, rv, floor)
            if arb is not None and stats[k + "@rms_floor"]["max"] > 1e-3:
                n_bad, worst, good_max, spread = _arbitrate(k, cv, rv, floor, arb)
                stats[k + "@rms_floor"].update(rays_arbitrated=n_bad, max_vs_nearest_reference_on_them=worst,
                                               references_apart_on_them=spread, max_before_arbitration=stats[k + "@rms_floor"]["max"])
                _record(name, stats)
                assert worst <= 1e-3, f"{name}/{k}: {n_bad} rays deviate from every reference evaluation ({worst:.3e}; references apart {spread:.3e})"
                stats[k + "@rms_floor"]["max"] = max(good_max, worst)
    _record(name, stats)
    for k, s in stats.items():
        if k in signed:
            continue            # asserted at the rms floor (its own entry); the strict-floor numbers are recorded
        assert s["max"] <= 1e-3, f"{name}/{k}: max rel err {s['max']:.3e} ({s['frac_beyond_1e-3']:.2e} of {s['n']} elements beyond 1e-3)"
    return stats


def test_rgb_full_frame_vs_reference_gpu():
    """BASELINE config 2, the whole frame: one fused launch vs the reference's staged chunk loop."""
    R, cand, ref = _models("rgb")
    for pose_k in (0, 7):
        ro, rd = _frame(pose_k)
        with torch.no_grad():
            want = R.render(ref, ro, rd, staged=True, perturb=False, bg_color=1)
            got = cand.render(ro, rd, staged=True, perturb=False, bg_color=1, image_width=W)
        assert set(k for k, v in want.items() if torch.is_tensor(v)) == {"image", "depth", "weights_sum"}
        _check(f"config2_rgb_pose{pose_k}", got, want)


def test_mask_full_frame_and_train_batch_vs_reference_gpu():
    """BASELINE config 4: (i) the whole frame with the object head, (ii) the incoherent train-style batch."""
    R, cand, ref = _models("mask")
    ro, rd = _frame(3)
    with torch.no_grad():
        want = R.render(ref, ro, rd, staged=True, perturb=False, bg_color=1, return_mask=1)
        got = cand.render(ro, rd, staged=True, perturb=False, bg_color=1, return_mask=1)
    assert "instance_mask_logits" in want and want["instance_mask_logits"].shape == (H * W, 2)
    _check("config4_mask_frame", got, want, signed=("instance_mask_logits",), arb=("mask", cand, ro, rd, dict(return_mask=1), R, ref))
    del want, got
    # (ii) 6000 uniformly random pixels across 24 seeded poses, then 4 local 8x8 patches (global rays first)
    g = torch.Generator().manual_seed(17)
    ros, rds = [], []
    per_pose = 6000 // 24
    for k in range(24):
        o, d = _frame(k)
        sel = torch.randint(0, H * W, (per_pose,), generator=g).to(DEV)
        ros.append(o[sel])
        rds.append(d[sel])
    for k in range(4):
        o, d = _frame(5 * k + 1)
        r0, c0 = int(torch.randint(0, H - 8, (1,), generator=g)), int(torch.randint(0, W - 8, (1,), generator=g))
        idx = (torch.arange(r0, r0 + 8)[:, None] * W + torch.arange(c0, c0 + 8)[None, :]).reshape(-1).to(DEV)
        ros.append(o[idx])
        rds.append(d[idx])
    ro, rd = torch.cat(ros).contiguous(), torch.cat(rds).contiguous()
    assert ro.shape[0] == 6000 + 4 * 64
    with torch.no_grad():
        # the trainer's call shape (trainer.py:407-409): non-staged, update_proposal=False, return_mask=1
        want = R.render(ref, ro, rd, staged=False, perturb=False, bg_color=1, update_proposal=False, return_mask=1)
        got = cand.render(ro, rd, staged=False, perturb=False, bg_color=1, update_proposal=False, return_mask=1)
    _check("config4_mask_train_batch", got, want, signed=("instance_mask_logits",), arb=("mask", cand, ro, rd, dict(return_mask=1), R, ref))
    s = STATS["config4_mask_frame"]["instance_mask_logits@rms_floor"], STATS["config4_mask_train_batch"]["instance_mask_logits@rms_floor"]
    assert max(x["max"] for x in s) <= 1e-3, s


def test_sam_full_frame_vs_reference_gpu():
    """BASELINE config 3: the reference in 160 non-staged calls of 5 rows (H=5, W=800), the candidate in ONE call."""
    R, cand, ref = _models("sam")
    ro, rd = _frame(11)
    with torch.no_grad():
        want = R.render_features_by_rows(ref, ro, rd, W, rows_per_call=5, perturb=False, bg_color=1)
        got = cand.render(ro, rd, staged=False, perturb=False, bg_color=1, return_feats=1, H=H, W=W, image_width=W)
    assert want["samvit"].shape == (H * W, 256) and got["samvit"].shape == (H, W, 256)
    stats = _check("config3_sam_frame", got, want, signed=("samvit",), arb=("sam", cand, ro, rd, dict(return_feats=1), R, ref))
    assert stats["samvit@rms_floor"]["max"] <= 1e-3, stats["samvit@rms_floor"]


def test_reference_gpu_vs_cpu_oracle_floors():
    """What can fp32 itself promise for the signed feature vectors?  The reference's own GPU path (cuBLAS fp32 SGEMM, CUB scans)
    against the fp32 CPU oracle (the same algorithm, serial association), on 4096 rays, at the strict SURVEY floor (1e-3) and at
    the rms floor.  The GPU oracle must sit well inside the rms floor; its strict-floor score is recorded next to the
    candidate's (gpurun_out/ref_gpu_parity.json) -- if the reference itself misses 1e-3 at the strict floor, that floor is not
    a property of the algorithm in fp32 and the rms floor stands."""
    import bench
    from oracle import render_oracle
    specs = render_oracle.default_specs(2)
    for wl, key in (("sam", "samvit"), ("mask", "instance_mask_logits")):
        R, cand, ref = _models(wl)
        ro, rd = _frame(2)
        lo = (H // 2) * W
        ro, rd = ro[lo:lo + 4096].contiguous(), rd[lo:lo + 4096].contiguous()
        kw = dict(return_feats=1, H=1, W=4096) if wl == "sam" else dict(return_mask=1)
        with torch.no_grad():
            gpu = R.render(ref, ro, rd, staged=False, perturb=False, bg_color=1, **kw)
            mine = cand.render(ro, rd, staged=False, perturb=False, bg_color=1, **kw)
        params = {k: v.detach().cpu() for k, v in cand.state_dict().items()}
        cpu, _ = render_oracle.run(params, specs, bench.default_opt(wl), ro.cpu(), rd.cpu(), bg_color=1, **kw)
        want = cpu[key].to(DEV)
        floor = _rms_floor(want)
        stats = {"reference_gpu@strict": _errors(gpu[key].reshape(want.shape), want, 1e-3),
                 "reference_gpu@rms_floor": _errors(gpu[key].reshape(want.shape), want, floor),
                 "candidate@strict": _errors(mine[key].reshape(want.shape), want, 1e-3),
                 "candidate@rms_floor": _errors(mine[key].reshape(want.shape), want, floor)}
        for k in ("image", "depth", "weights_sum"):
            stats["reference_gpu/" + k] = _errors(gpu[k], cpu[k].to(DEV), 1e-3)
            stats["candidate/" + k] = _errors(mine[k], cpu[k].to(DEV), 1e-3)
        _record(f"floors_{wl}", stats)
        assert stats["reference_gpu@rms_floor"]["max"] <= 1e-3 and stats["candidate@rms_floor"]["max"] <= 1e-3, stats
        for k in ("image", "depth", "weights_sum"):
            assert stats["candidate/" + k]["max"] <= 1e-3, (k, stats["candidate/" + k])


def test_perturbed_staged_render_vs_reference_gpu_same_seed():
    """perturb=True through the staged loop (the SAM stage renders its RGB frames this way, trainer.py:513-514): the fused
    kernel draws the jitter with torch.rand in the reference's order and chunking, so with the same torch seed both sides see
    the same random numbers."""
    R, cand, ref = _models("rgb")
    ref.opt.max_ray_batch = cand.opt.max_ray_batch = 4096
    ro, rd = _frame(4)
    lo = 300 * W
    ro, rd = ro[lo:lo + 20000].contiguous(), rd[lo:lo + 20000].contiguous()      # five chunks, the last one ragged
    with torch.no_grad():
        torch.manual_seed(21)
        want = R.render(ref, ro, rd, staged=True, perturb=True, bg_color=1)
        torch.manual_seed(21)
        got = cand.render(ro, rd, staged=True, perturb=True, bg_color=1)
        plain = cand.render(ro, rd, staged=True, perturb=False, bg_color=1)
    assert not torch.equal(got["depth"], plain["depth"])
    _check("perturb_rgb_staged", got, want)
