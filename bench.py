#!/usr/bin/env python
"""bench.py -- Mrays/s of the SANeRF-HQ render hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rgb|sam|mask] [--only] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full-frame render (BASELINE config #2: synthetic Lego-style 800x800 RGB, hashgrid L=16 T=2^19, MLP 2x64,
samples 128+64+32) of a different pose of a seeded 24-pose orbit, through this repo's `NeRFNetwork.render` = one persistent
launch of the fused sm_100a kernel (+ one tensor-core head launch for the SAM-feature / object workloads).

The JSON line (rank 0) is the headline workload (`--workload`, default rgb = config #2); unless `--only` is given it also
carries `workloads: {sam: {...}, mask: {...}}` = BASELINE configs #3 / #4 measured the same way, and at N = 8 (and N = 1) `config5` = the
1600x1600 RGB+SAM frame sharded `rank r <- rows [200r, 200r+200)` with a bit-for-bit check against the 1-GPU frame.

Per workload: value = whole-job Mrays/s with rays resident in HBM; e2e = the same through the public API with HOST (pinned)
rays: H2D of the rays and D2H of EVERY requested output inside the timed region; kernels = per-kernel live CUDA-event times
with the roofline that bounds each (render kernel: algorithmic gather + I/O bytes vs measured HBM copy peak; heads: useful
MLP flops vs measured sustained bf16 tensor peak, also counting the 3 split-precision products); ref_gpu_baseline = the
REFERENCE ITSELF on the same GPU (its unmodified Python, byte-compiled, on its own CUDA kernels compiled verbatim:
oracle/_ref) at max_ray_batch 4096 and 16384, same weights and poses, with the candidate's worst relative deviation from it;
cpu_baseline = the reference's CPU path on the host cores over a bounded sample.

Weak scaling at N GPUs: a step renders N views of the orbit, one 800x800 frame per GPU (model replicated, no data-path
collective), and the N composited frames are left on every rank through NVLink peer memory (sanerf_hq_b200/parallel.py:
in-kernel peer stores + copy-engine pushes + one flag barrier; in-place NCCL all-gathers as the fallback) inside the timed
region.  (config5 is the strong-scaling counterpart: ONE frame, its rows sharded over the ranks.)

--impl reference times the reference's own CPU implementation on the host cores: the reference's unmodified renderer / network
Python (oracle/_ref/bytecode) over the C restatement of its two CUDA-only encoder kernels (the reference has no CPU encoder), all
host threads, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time
import warnings

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

# algorithmic bytes per ray: gathers = samples x levels x 8 corners x C x 4 B, + ray in / pixel out
GATHER_RGB = 40960 + 20480 + 32768          # prop0 + prop1 + grid (C=2)
GATHER_FEAT = 131072                        # s_grid or m_grid: 32 samples x 16 levels x 8 corners x 8 channels x 4 B
BYTES_PER_RAY = {"rgb": 94252, "sam": 226348, "mask": 225332}
FLOPS_PER_RAY = {"rgb": 530560, "sam": 1221760, "mask": 7100544}
HEAD_FLOPS_PER_RAY = {"samvit_mlp_kernel": 691200, "mask_head_kernel": 6569984}
H_FRAME, W_FRAME, N_POSES = 800, 800, 24
METRIC = "Mrays/sec (RGB+feat render, 800x800)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="rgb", choices=["rgb", "sam", "mask"])
    ap.add_argument("--only", action="store_true", help="measure only --workload (no workloads block, no reference GPU baseline)")
    ap.add_argument("--impl", default="candidate", choices=["candidate", "reference"])
    ap.add_argument("--height", type=int, default=H_FRAME)
    ap.add_argument("--width", type=int, default=W_FRAME)
    ap.add_argument("--groups", type=int, default=0, help="row groups per frame for the SAM workload at N > 1 (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="N = 1: skip the single-GPU 1600x1600 frame of BASELINE config 5")
    return ap.parse_args()


def default_opt(workload):
    """The opt fields the hot path reads, as main.py produces them for the shipped scripts."""
    from types import SimpleNamespace
    return SimpleNamespace(bound=128, contract=True, min_near=0.2, density_thresh=10, render_mesh=False,
                           num_steps=[128, 64, 32], background="last_sample", with_sam=workload == "sam",
                           with_mask=workload == "mask", mask_mlp_type="default", sam_use_view_direction=True, n_inst=2,
                           max_ray_batch=4096, lambda_proposal=1, lambda_distort=0.02)


def build_model(workload, device):
    """Random-init weights of the reference architecture: seed-0 constructor (nn.Linear Kaiming init in the
    reference's construction order), hash tables re-drawn U(-1,1) with seeds 1+k."""
    from sanerf_hq_b200.network import NeRFNetwork
    torch.manual_seed(0)
    model = NeRFNetwork(default_opt(workload))
    k = 0
    for name, p in model.state_dict().items():
        if name.endswith("embeddings"):
            g = torch.Generator().manual_seed(1 + k)
            p.copy_(torch.rand(p.shape, generator=g) * 2 - 1)
            k += 1
    return model.eval().to(device)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 50 ms while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------------------------------
_CPU_REF = {}


def cpu_baseline(workload, model_sd, n_chunks, H, W, threads=None):
    """The reference's CPU path over `n_chunks` x 4096 rays of pose 0 (rows from the middle of the frame), all host threads.
    kind "reference": the reference's own renderer.py / network.py (oracle/_ref/bytecode, byte-compiled unmodified) with the C
    restatement of its two CUDA-only encoder kernels behind its `_gridencoder` / `_shencoder` imports; kind "port": the
    oracle's torch restatement of the same Python (when oracle/_ref/bytecode is not staged).
    Returns (Mrays/s, cores, kind, sample description, seconds)."""
    from oracle import kernels as K
    from oracle import ref_runtime as R
    from oracle import render_oracle as O
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    K.set_num_threads(cores)
    params = {k: v.detach().cpu() for k, v in model_sd.items()}
    first_row = max(0, H // 2 - (n_chunks * 4096 // W) // 2)
    rows = (first_row, min(H, first_row + -(-n_chunks * 4096 // W)))
    rays_o, rays_d = get_rays(orbit_pose(0), lego_intrinsics(H, W), H, W, rows=rows)
    n = min(rays_o.shape[0], n_chunks * 4096)
    if R.available("cpu"):
        kind = "reference"
        if workload not in _CPU_REF:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _CPU_REF[workload] = R.build_network(default_opt(workload), params, device="cpu", backend="cpu")
        model = _CPU_REF[workload]

        def run(o, d, **kw):
            with torch.no_grad():
                return R.render(model, o, d, backend="cpu", staged=False, perturb=False, bg_color=1, **kw)
    else:
        kind = "port"
        opt = O.default_opt(with_sam=workload == "sam", with_mask=workload == "mask")
        specs = O.default_specs(2)

        def run(o, d, **kw):
            return O.run(params, specs, opt, o, d, **kw)
    run(rays_o[:512], rays_d[:512], **({"return_mask": 1} if workload == "mask" else {}))  # warm
    t0 = time.perf_counter()
    for head in range(0, n, 4096):
        m = min(4096, n - head)
        kw = {}
        if workload == "sam":
            kw = dict(return_feats=1, H=1, W=m)
        elif workload == "mask":
            kw = dict(return_mask=1)
        run(rays_o[head:head + m], rays_d[head:head + m], **kw)
    dt = time.perf_counter() - t0
    return n / dt / 1e6, cores, kind, f"{n} rays of pose 0 (rows {rows[0]}..{rows[1]}) in {-(-n // 4096)} chunks of 4096, {dt:.1f} s", dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on the host cores, rank 0 only."""
    if rank != 0:
        return
    torch.manual_seed(0)
    model = build_model(args.workload, "cpu")
    sd = model.state_dict()
    chunks_per_step = 8
    for _ in range(args.warmup):
        cpu_baseline(args.workload, sd, 1, args.height, args.width)
    t_total, rays_total, cores, sample, kind = 0.0, 0, 1, "", "port"
    for _ in range(args.steps):
        mr, cores, kind, sample, dt = cpu_baseline(args.workload, sd, chunks_per_step, args.height, args.width)
        t_total += dt
        rays_total += int(round(mr * 1e6 * dt))
    value = rays_total / t_total / 1e6
    what = ("the reference's own renderer.py / network.py (bytecode, oracle/_ref/bytecode) + C restatement of its two CUDA-only encoder "
            "kernels" if kind == "reference" else "oracle port of the reference's renderer / network + C restatement of its encoder kernels")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} {args.height}x{args.width} per GPU (global frame {args.height}x{args.width}), hashgrid L=16 "
                                   f"T=2^19, MLP 2x64, samples 128+64+32, reference algorithm in chunks of 4096 rays on the host CPU",
                       "rays_per_step": chunks_per_step * 4096, "device": "host CPU", "what": what,
                       "weights": "random init: seed-0 constructor, hash tables U(-1,1)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": f"per step: {sample}"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# reference GPU baseline (R-GPU): the reference's own render on its own CUDA kernels, same GPU, same weights, same poses
# ---------------------------------------------------------------------------------------------------------------------------
def ref_gpu_baseline(wl, model, dev, H, W, poses, intr, cand_out0):
    """Mrays/s of the reference's `render(staged=True)` (SAM: 5-row non-staged chunks, the only way its API renders features at
    800x800) at max_ray_batch 4096 (BASELINE) and 16384 (its default), CUDA events, 1 warm-up + 2 timed
    frames each; and the candidate's worst relative deviation from the reference's frame of pose 0 (SURVEY.md 8d tolerance)."""
    from oracle import ref_runtime as R
    from sanerf_hq_b200.rays import get_rays
    if not R.available("cuda"):
        return {"unavailable": "oracle/_ref is not staged (python oracle/stage_ref.py; python oracle/build_ref.py)"}
    torch.backends.cuda.matmul.allow_tf32 = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = R.build_network(default_opt(wl), model.state_dict(), device=dev)
    rays = [tuple(t.to(dev) for t in get_rays(poses[k], intr, H, W)) for k in range(3)]      # the candidate's rays, bit for bit
    n = H * W
    out = {"what": "reference nerf/renderer.py + network.py (unmodified, byte-compiled) on the reference's CUDA kernels compiled verbatim "
                   "(oracle/_ref), fp32, TF32 off, same weights / poses / GPU", "unit": "Mrays/s"}

    def frame(k):
        ro, rd = rays[k % 3]
        if wl == "sam":
            return R.render_features_by_rows(ref, ro, rd, W, rows_per_call=5, perturb=False, bg_color=1)
        return R.render(ref, ro, rd, staged=True, perturb=False, bg_color=1, **({"return_mask": 1} if wl == "mask" else {}))

    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        first = None
        for batch in ((4096,) if wl == "sam" else (4096, 16384)):
            ref.opt.max_ray_batch = batch
            r = frame(0)
            first = first or r
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for k in (1, 2):
                frame(k)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 2
            key = "max_ray_batch_4096" if batch == 4096 else "max_ray_batch_16384"
            if wl == "sam":
                key = "rows_5_per_call_4000_rays"
            out[key] = {"value": n / (ms * 1e-3) / 1e6, "ms_per_frame": ms}
        # parity of the candidate's pose-0 frame against the reference's (same rays)
        dev_max, dev_frac = {}, {}
        for k, rv in first.items():
            if torch.is_tensor(rv) and k in cand_out0:
                cv = cand_out0[k].reshape(rv.shape)
                floor = 1e-3
                if k in ("samvit", "instance_mask_logits"):
                    floor = max(1e-3, 0.1 * float(rv.double().pow(2).mean().sqrt()))
                worst, beyond = 0.0, 0
                cf, rf = cv.reshape(-1), rv.reshape(-1)
                for head in range(0, rf.numel(), 1 << 24):
                    x, y = cf[head:head + (1 << 24)].double(), rf[head:head + (1 << 24)].double()
                    e = (x - y).abs() / y.abs().clamp(min=floor)
                    worst = max(worst, float(e.max()))
                    beyond += int((e > 1e-3).sum())
                dev_max[k] = worst
                dev_frac[k] = beyond / max(1, rf.numel())
        out["candidate_max_rel_dev_pose0"] = dev_max
        out["candidate_frac_beyond_tolerance_pose0"] = dev_frac
        out["tolerance"] = ("|cand - ref| <= 1e-3 * max(|ref|, floor); floor 1e-3 (samvit / logits: max(1e-3, 0.1 rms)).  A few rays per frame "
                            "(near-ties in sample_pdf, amplified by the head MLPs) exceed it between ANY two fp32 evaluations of the reference "
                            "algorithm -- the reference GPU against itself at another batch shape, or against the CPU oracle: "
                            "tests/test_ref_gpu_frames.py arbitrates those rays")
    del ref
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# candidate arm
# ---------------------------------------------------------------------------------------------------------------------------
def result_spec(wl):
    spec = {"image": (3,), "depth": (), "weights_sum": ()}
    if wl == "sam":
        spec["samvit"] = (256,)
    if wl == "mask":
        spec["instance_mask_logits"] = (2,)
    return spec


def measure(wl, args, ctx, H, W, headline, rows_of_rank=None, global_h=None, steps=None, warmup=None, tag=None, baselines=True):
    """All numbers of one workload on this rank's `H x W` rays.  Default (weak scaling): every rank renders its own view of the
    orbit, H x W pixels; with `rows_of_rank` / `global_h` (strong scaling): rows `rows_of_rank` of ONE `global_h` x W frame.
    Returns a dict on rank 0."""
    import torch.distributed as dist
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.parallel import FrameGather
    from sanerf_hq_b200.rays import get_rays
    rank, world, dev = ctx["rank"], ctx["world"], ctx["dev"]
    steps, warmup = steps or args.steps, warmup or args.warmup
    from sanerf_hq_b200.rays import lego_intrinsics, orbit_pose
    strong = rows_of_rank is not None
    global_h = global_h if strong else H
    rows = rows_of_rank if strong else (0, H)
    intr = lego_intrinsics(global_h, W)
    # weak scaling: view (i * world + rank) of the orbit in step i; strong scaling: every rank works on view i
    poses = [orbit_pose(k if strong else (k * world + rank) % N_POSES, N_POSES) for k in range(N_POSES)]
    model = build_model(wl, dev)
    n_local, n_total = H * W, H * W * world
    spec = result_spec(wl)
    keys = list(spec)
    kw = dict(return_mask=1) if wl == "mask" else (dict(return_feats=1) if wl == "sam" else {})
    groups = 1
    if wl == "sam" and world > 1:
        groups = args.groups or 8
    fg = FrameGather(n_local, spec, dev)

    n_res = min(N_POSES, steps + warmup)
    # (strong scaling: the rows are sliced out of the full frame's rays so that every rank sees bit-identical ray values)
    lo_ray, hi_ray = rows[0] * W, rows[1] * W
    host_rays = [tuple(t[lo_ray:hi_ray].contiguous().pin_memory() for t in get_rays(poses[k], intr, global_h, W)) for k in range(n_res)]
    dev_rays = [(o.to(dev), d.to(dev)) for o, d in host_rays]
    hint = W if (n_local % (4 * W) == 0 and W % 4 == 0) else None

    def render_frame(ro, rd):
        return fg.render(model, ro, rd, groups=groups, perturb=False, image_width=hint, **kw)

    flush = None if args.no_l2_flush else ctx["flush"]

    def timed_loop(step_fn, n_steps, n_warm, finish_fn=None):
        for i in range(n_warm):
            step_fn(i)
        if finish_fn is not None:
            finish_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        n0 = _lib.launch_counter["n"]
        for i in range(n_steps):
            if flush is not None:
                flush.fill_(i & 0xFF)          # evict L2 between timed iterations (outside the per-step events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(n_warm + i)
            if finish_fn is not None and i == n_steps - 1:
                finish_fn()                    # drain what the last step left in flight: inside the timed region
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if os.environ.get("SANERF_BENCH_DEBUG"):
            print(f"[rank {rank}] {wl} step ms: " + " ".join(f"{a.elapsed_time(b):.2f}" for a, b in evs), file=sys.stderr, flush=True)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_counter["n"] - n0

    res = {}
    with torch.no_grad():
        # ---- device-resident arm ("value") ------------------------------------------------------
        sink = {}

        def step_dev(i):
            ro, rd = dev_rays[i % n_res]
            sink["out"] = render_frame(ro, rd)

        sampler = ClockSampler(ctx["local_rank"]) if headline else None
        if sampler:
            sampler.start()
        total_ms, launches = timed_loop(step_dev, steps, warmup)
        if sampler:
            sampler.stop_flag = True
        ms_per_step = total_ms / steps
        value = n_total / (ms_per_step * 1e-3) / 1e6

        # ---- per-kernel times (same call, same tile hint, L2 flushed) for the rooflines -----------------------------------
        _lib.kernel_events = []
        reps = max(3, min(10, steps))
        for i in range(reps):
            ro, rd = dev_rays[i % n_res]
            if flush is not None:
                flush.fill_(i)
            render_frame(ro, rd)
        torch.cuda.synchronize()
        per_kernel = {}
        for label, a, b in _lib.kernel_events:
            per_kernel.setdefault(label, 0.0)
            per_kernel[label] += a.elapsed_time(b) / reps
        _lib.kernel_events = None

        # ---- end-to-end arm: host rays -> H2D -> render (public API) -> D2H of EVERY requested output -----------------------
        # rank 0 reads the whole (gathered) frame back, like the reference's rank-0 image writer; the other ranks read nothing
        host_out = {k: torch.empty((n_total,) + s).pin_memory() for k, s in spec.items()} if rank == 0 else None
        o_dev, d_dev = torch.empty(n_local, 3, device=dev), torch.empty(n_local, 3, device=dev)

        skip = os.environ.get("SANERF_E2E_SKIP", "")      # debugging aid: "h2d" / "d2h" leave that copy out

        def step_e2e(i):
            ho, hd = host_rays[i % n_res]
            if "h2d" not in skip:
                o_dev.copy_(ho, non_blocking=True)
                d_dev.copy_(hd, non_blocking=True)
            out = render_frame(o_dev, d_dev)
            if host_out is not None and "d2h" not in skip:
                for k in keys:
                    host_out[k].copy_(out[k], non_blocking=True)

        e2e_ms = timed_loop(step_e2e, steps, warmup)[0] / steps
        e2e_value = n_total / (e2e_ms * 1e-3) / 1e6

        # ---- the same, streamed: the H2D of frame i+1 and the D2H of frame i-1 run on a copy stream while frame i renders
        # (double-buffered ray buffers; FrameGather double-buffers the results).  Every copy is waited for by the main stream
        # inside a timed step, the last D2H inside the last one.
        copy_stream = torch.cuda.Stream(device=dev)
        ray_buf = [(torch.empty(n_local, 3, device=dev), torch.empty(n_local, 3, device=dev)) for _ in range(2)]
        h2d_done, rays_free, d2h_done = [None, None], [None, None], [None]

        def prefetch(i):
            slot = i % 2
            ho, hd = host_rays[i % n_res]
            with torch.cuda.stream(copy_stream):
                if rays_free[slot] is not None:
                    copy_stream.wait_event(rays_free[slot])
                ray_buf[slot][0].copy_(ho, non_blocking=True)
                ray_buf[slot][1].copy_(hd, non_blocking=True)
                h2d_done[slot] = torch.cuda.Event()
                h2d_done[slot].record(copy_stream)

        def step_stream(i):
            slot, main = i % 2, torch.cuda.current_stream(dev)
            if h2d_done[slot] is None:
                prefetch(i)
            main.wait_event(h2d_done[slot])
            h2d_done[slot] = None
            prefetch(i + 1)
            out = render_frame(*ray_buf[slot])
            rays_free[slot] = torch.cuda.Event()
            rays_free[slot].record(main)
            if d2h_done[0] is not None:
                main.wait_event(d2h_done[0])          # the previous frame has left the device before this step ends
            if host_out is not None:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(rays_free[slot])
                    for k in keys:
                        host_out[k].copy_(out[k], non_blocking=True)
                    d2h_done[0] = torch.cuda.Event()
                    d2h_done[0].record(copy_stream)

        def drain():
            if d2h_done[0] is not None:
                torch.cuda.current_stream(dev).wait_event(d2h_done[0])

        stream_ms = timed_loop(step_stream, steps, warmup, finish_fn=drain)[0] / steps
        stream_value = n_total / (stream_ms * 1e-3) / 1e6

        # ---- informational: the same frame through render_image: host pose in (64 B), 8-bit image out -------
        cam_value = None
        if wl == "rgb" and world == 1 and headline:
            u8_host = torch.empty(n_local, 3, dtype=torch.uint8).pin_memory()

            def step_cam(i):
                out = model.render_image(poses[i % N_POSES], intr, H, W, return_uint8=True)
                u8_host.copy_(out["image_u8"], non_blocking=True)

            cam_ms = timed_loop(step_cam, steps, warmup)[0] / steps
            cam_value = n_local / (cam_ms * 1e-3) / 1e6

        # ---- informational: the SAM feature frame in the shape its consumer wants (nerf/trainer.py:540-546 resizes
        # the [H,W,256] frame to [1,256,64,64] for the SAM decoder): host rays in, image + depth + weights_sum + the 64x64 NCHW
        # feature tensor out -- the permute and the bilinear resize happen on the device, 4 MB instead of 655 MB cross PCIe
        consumer_value = None
        if wl == "sam" and world == 1 and not strong:
            try:
                small = {"image": torch.empty(n_local, 3).pin_memory(), "depth": torch.empty(n_local).pin_memory(),
                         "weights_sum": torch.empty(n_local).pin_memory(), "samvit_nchw": torch.empty(1, 256, 64, 64).pin_memory()}

                def step_consumer(i):
                    ho, hd = host_rays[i % n_res]
                    o_dev.copy_(ho, non_blocking=True)
                    d_dev.copy_(hd, non_blocking=True)
                    out = model.render(o_dev, d_dev, staged=False, perturb=False, return_feats=1, H=H, W=W, image_width=hint,
                                       feature_layout="nchw", feature_size=(64, 64))
                    for k, t in small.items():
                        t.copy_(out[k], non_blocking=True)

                consumer_ms = timed_loop(step_consumer, steps, warmup)[0] / steps
                consumer_value = n_local / (consumer_ms * 1e-3) / 1e6
            except Exception as e:   # noqa: BLE001 -- informational figure only
                print(f"bench.py: consumer-shaped SAM frame not measured ({type(e).__name__}: {e})", file=sys.stderr)

        # pose-0 frame of this rank for the parity figure / checksum
        out0 = {k: v.clone() for k, v in render_frame(*dev_rays[0]).items()}
        torch.cuda.synchronize()

    if rank == 0:
        peaks = ctx["peaks"]
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        tc_peak = peaks.get("bf16_tflops_sustained", 1399.0)
        traffic = ctx["traffic"]
        kernels = []
        for label, k_ms in per_kernel.items():
            if label == "render_kernel":
                bpr = GATHER_RGB + 44 + (GATHER_FEAT + 163 * 4 if wl == "sam" else 0) + (18 * 4 * 32 + 32 * 4 if wl == "mask" else 0)
                ach = bpr * n_local / (k_ms * 1e-3) / 1e9
                kernels.append({"kernel": "sanerf::render_kernel", "ms": k_ms, "bound": "hbm", "algorithmic_bytes_per_ray": bpr,
                                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                "traffic": (traffic.get(wl + "_render") or (traffic.get(wl) if wl == "rgb" else None))})
            else:
                fl = HEAD_FLOPS_PER_RAY[label]
                ach = fl * n_local / (k_ms * 1e-3) / 1e12
                bpr = (163 * 4 + 1024) if label == "samvit_mlp_kernel" else (GATHER_FEAT + 18 * 4 * 32 + 32 * 4 + 8)
                kernels.append({"kernel": "sanerf::" + label, "ms": k_ms, "bound": "tensor", "flops_per_ray": fl, "achieved": ach,
                                "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "frac_counting_3_split_products": 3 * ach / tc_peak,
                                "algorithmic_bytes_per_ray": bpr, "hbm_convention_gbs": bpr * n_local / (k_ms * 1e-3) / 1e9,
                                "traffic": traffic.get(wl + "_head")})
        dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
        step_ms_kernels = sum(k["ms"] for k in kernels)
        achieved = BYTES_PER_RAY[wl] * n_local / (step_ms_kernels * 1e-3) / 1e9 if kernels else None
        res = {
            "value": value, "unit": "Mrays/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "config": {"workload": f"{tag or wl} {H}x{W} per GPU (" + (f"rows of one {global_h}x{W} frame" if strong else
                                                                      f"{world} view(s) of {H}x{W} per step") +
                                   "), hashgrid L=16 T=2^19, MLP 2x64, samples 128+64+32, one fused render launch per frame "
                                   "(reference: 4096 rays/batch)",
                       "rays_per_step": n_total, "poses": n_res,
                       "parallelism": (f"rows of one frame sharded x{world}" if strong else f"one view per GPU x{world}") +
                                      f", results gathered on every rank, transport {fg.transport}" + (f", {groups} row groups" if groups > 1 else ""),
                       "l2": "no flush" if flush is None else "L2 flushed between timed steps (256 MiB fill, outside the step events)",
                       "weights": "random init: seed-0 constructor, hash tables U(-1,1)"},
            # whole-path roofline by the SURVEY 8d convention (all kernels of the step), then the per-kernel list
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak if achieved else None,
                         "traffic": traffic.get(wl), "peak_source": peak_src,
                         "kernel": " + ".join(k["kernel"] for k in kernels), "kernel_ms": step_ms_kernels,
                         "dominant_kernel": dom["kernel"] if dom else None,
                         "algorithmic_bytes_per_ray": BYTES_PER_RAY[wl],
                         "mlp_tflops": FLOPS_PER_RAY[wl] * n_local / (step_ms_kernels * 1e-3) / 1e12 if kernels else None},
            "kernels": kernels,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * n_local * 12 * world,
                    "d2h_bytes_per_step": int(sum(t.numel() * 4 for t in host_out.values())),
                    "d2h": "rank 0 reads every gathered output: " + ", ".join(keys), "render_image_uint8_value": cam_value,
                    "sam_consumer_nchw64_value": consumer_value,
                    "streamed_value": stream_value, "streamed_ms_per_step": stream_ms,
                    "streamed": "same copies, on a copy stream: H2D of the next and D2H of the previous frame overlap the render"},
            "gpu_launches": int(launches),
            "checksum_pose0": {k: float(v.double().sum()) for k, v in out0.items()},
        }
        if sampler:
            res["clocks"] = sampler.summary()
        if world == 1 and baselines and not args.no_cpu_baseline:
            mr, cores, kind, sample, _ = cpu_baseline(wl, model.state_dict(), (40 if wl == "rgb" else 12) if headline else 8, H, W)
            res["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample}
        if world == 1 and baselines and not args.no_ref_gpu and not args.only:
            try:
                res["ref_gpu_baseline"] = ref_gpu_baseline(wl, model, dev, H, W, poses, intr, out0)
                best = res["ref_gpu_baseline"].get("max_ray_batch_4096") or res["ref_gpu_baseline"].get("rows_5_per_call_4000_rays")
                if best:
                    res["ref_gpu_baseline"]["speedup_vs_4096"] = value / best["value"]
            except Exception as e:   # the baseline is a report, never a reason to lose the candidate's numbers
                res["ref_gpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"}
    fg.close()
    ctx["last_out0"] = out0
    ctx["last_model"] = model
    return res


def config5(args, ctx):
    """BASELINE config #5: 1600x1600 RGB + SAM feature, rank r <- rows [200r, 200r+200) at 8 GPUs (H/world rows in
    general), outputs gathered on every rank; the gathered frame must equal the 1-GPU render of the same frame bit for bit
    (rank 0 renders the whole frame once, untimed)."""
    import torch.distributed as dist
    rank, world, dev = ctx["rank"], ctx["world"], ctx["dev"]
    Hg = Wg = 1600
    rows = Hg // world
    res = measure("sam", args, ctx, rows, Wg, headline=False, rows_of_rank=(rank * rows, (rank + 1) * rows), global_h=Hg,
                  steps=min(args.steps, 10), warmup=3, tag="config5 sam (strong-sharded 1600x1600)", baselines=False)
    model, got = ctx.pop("last_model"), ctx.pop("last_out0")
    equal = {}
    if rank == 0:
        from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
        with torch.no_grad():
            ro, rd = (t.to(dev) for t in get_rays(orbit_pose(0, N_POSES), lego_intrinsics(Hg, Wg), Hg, Wg))   # as in measure()
            want = model.render(ro, rd, staged=False, perturb=False, return_feats=1, H=Hg, W=Wg, image_width=Wg)
            equal = {k: bool(torch.equal(got[k].reshape(-1), want[k].reshape(-1))) for k in got}
        res["scaling"] = "strong"
        res["gathered_equals_single_gpu_bit_for_bit"] = equal
    if world > 1:
        dist.barrier()
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the candidate arm has no CPU path; use --impl reference for the CPU arm)")

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks, traffic = {}, {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))   # from the committed ncu captures
    except Exception:
        pass
    ctx = {"rank": rank, "local_rank": local_rank, "world": world, "dev": dev, "peaks": peaks, "traffic": traffic,
           "flush": torch.empty(256 << 20, dtype=torch.uint8, device=dev)}
    H, W = args.height, args.width

    head = measure(args.workload, args, ctx, H, W, headline=True)
    ctx.pop("last_model", None), ctx.pop("last_out0", None)
    torch.cuda.empty_cache()
    extra = {}
    if not args.only:
        quick = dict(steps=max(3, min(args.steps, 10)), warmup=3)
        for wl in ("rgb", "sam", "mask"):
            if wl != args.workload:
                extra[wl] = measure(wl, args, ctx, H, W, headline=False, **quick)
                ctx.pop("last_model", None), ctx.pop("last_out0", None)
                torch.cuda.empty_cache()
        if world == 8 and H == H_FRAME and W == W_FRAME:
            extra["config5"] = config5(args, ctx)
        elif world == 1 and H == H_FRAME and W == W_FRAME and not args.no_config5:
            # the same 1600x1600 frame on ONE GPU: the denominator of config 5's strong scaling.  No collective is involved at
            # N = 1, so a failure here (e.g. a smaller GPU running out of memory) must not cost the headline line.
            try:
                extra["config5"] = config5(args, ctx)
            except Exception as e:   # noqa: BLE001
                extra["config5"] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic"}
        for k in ("config", "roofline", "kernels", "e2e", "gpu_launches", "clocks", "cpu_baseline", "ref_gpu_baseline", "checksum_pose0"):
            if k in head:
                line[k] = head[k]
        if extra:
            c5 = extra.pop("config5", None)
            line["workloads"] = extra
            if c5:
                line["config5"] = c5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
