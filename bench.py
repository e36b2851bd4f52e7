#!/usr/bin/env python
"""bench.py -- Mrays/s of the SANeRF-HQ render hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rgb|sam|mask] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full-frame render (BASELINE config #2: synthetic Lego-style 800x800 RGB, hashgrid L=16
T=2^19, MLP 2x64, samples 128+64+32) of a different pose of a seeded 24-pose orbit, through this
repo's `NeRFNetwork.render(staged=True)` = one persistent launch of the fused sm_100a kernel.

Weak scaling at N GPUs: the frame grows to (800*N) x 800 rows, rank r renders rows [800r, 800(r+1))
(model replicated, no data-path collective) and ONE all-gather leaves the composited frame on every
rank; the all-gather is inside the timed region.

JSON line (rank 0): value = whole-job Mrays/s with rays resident in HBM; e2e = same through the public
API with HOST (pinned) rays: H2D of the rays and D2H of the image inside the timed region; roofline =
algorithmic gather+IO bytes / kernel time vs measured HBM peak; cpu_baseline = the oracle port of the
reference's CPU path on the host cores over a bounded sample.

--impl reference times the reference's own algorithm on the host CPU (the reference is Python + CUDA-only
encoders, so the CPU arm is the oracle port: reference renderer/network restated in torch + C restatement
of the two CUDA-only encoder kernels), all host threads, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

# algorithmic bytes per ray: gathers = samples x levels x 8 corners x C x 4 B, + ray in / pixel out
BYTES_PER_RAY = {"rgb": 94252, "sam": 226348, "mask": 225332}
FLOPS_PER_RAY = {"rgb": 530560, "sam": 1221760, "mask": 7100544}
H_FRAME, W_FRAME, N_POSES = 800, 800, 24
TILE_HINT = os.environ.get("SANERF_BENCH_TILE_HINT", "1") != "0"   # pass image_width= (row-major frame) to the renderer


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="rgb", choices=["rgb", "sam", "mask"])
    ap.add_argument("--impl", default="candidate", choices=["candidate", "reference"])
    ap.add_argument("--height", type=int, default=H_FRAME)
    ap.add_argument("--width", type=int, default=W_FRAME)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
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


def cpu_baseline(workload, model_sd, n_chunks, H, W, threads=None):
    """Oracle port of the reference's CPU path, timed on the host cores over `n_chunks` x 4096 rays of pose 0
    (rows from the middle of the frame).  Returns (Mrays/s, cores, sample description, seconds)."""
    from oracle import kernels as K
    from oracle import render_oracle as O
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    K.set_num_threads(cores)
    opt = O.default_opt(with_sam=workload == "sam", with_mask=workload == "mask")
    specs = O.default_specs(2)
    params = {k: v.detach().cpu() for k, v in model_sd.items()}
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    first_row = max(0, H // 2 - (n_chunks * 4096 // W) // 2)
    rows = (first_row, min(H, first_row + -(-n_chunks * 4096 // W)))
    rays_o, rays_d = get_rays(orbit_pose(0), lego_intrinsics(H, W), H, W, rows=rows)
    n = min(rays_o.shape[0], n_chunks * 4096)
    kw = {}
    O.run(params, specs, opt, rays_o[:512], rays_d[:512], **({"return_mask": 1} if workload == "mask" else {}))  # warm
    t0 = time.perf_counter()
    for head in range(0, n, 4096):
        m = min(4096, n - head)
        if workload == "sam":
            kw = dict(return_feats=1, H=1, W=m)
        elif workload == "mask":
            kw = dict(return_mask=1)
        O.run(params, specs, opt, rays_o[head:head + m], rays_d[head:head + m], **kw)
    dt = time.perf_counter() - t0
    return n / dt / 1e6, cores, f"{n} rays of pose 0 (rows {rows[0]}..{rows[1]}) in {-(-n // 4096)} chunks of 4096, {dt:.1f} s", dt


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host CPU (oracle port), rank 0 only."""
    if rank != 0:
        return
    torch.manual_seed(0)
    model = build_model(args.workload, "cpu")
    sd = model.state_dict()
    chunks_per_step = 8
    for _ in range(args.warmup):
        cpu_baseline(args.workload, sd, 1, args.height, args.width)
    t_total, rays_total, cores, sample = 0.0, 0, 1, ""
    for _ in range(args.steps):
        mr, cores, sample, dt = cpu_baseline(args.workload, sd, chunks_per_step, args.height, args.width)
        t_total += dt
        rays_total += int(round(mr * 1e6 * dt))
    value = rays_total / t_total / 1e6
    line = {"impl": "reference", "metric": "Mrays/sec (RGB+feat render, 800x800)", "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} {args.height}x{args.width} per GPU (global frame {args.height}x{args.width}), hashgrid L=16 "
                                   f"T=2^19, MLP 2x64, samples 128+64+32, reference algorithm in chunks of 4096 rays on the host CPU",
                       "rays_per_step": chunks_per_step * 4096, "device": "host CPU",
                       "weights": "random init: seed-0 constructor, hash tables U(-1,1)"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"per step: {sample}"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    from sanerf_hq_b200 import _lib
    from sanerf_hq_b200.parallel import gather_dict
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H, W, wl = args.height, args.width, args.workload
    model = build_model(wl, dev)
    intr = lego_intrinsics(H * world, W)       # weak scaling: the global frame is (H*world) x W, this rank owns H rows
    rows = (rank * H, (rank + 1) * H)
    n_local = H * W
    n_total = n_local * world
    keys = ["image", "depth", "weights_sum"] + (["samvit"] if wl == "sam" else []) + (["instance_mask_logits"] if wl == "mask" else [])

    # rays of every pose of the orbit, resident in HBM before the timed region (and pinned on the host for e2e)
    poses = [orbit_pose(k, N_POSES) for k in range(N_POSES)]
    n_res = min(N_POSES, args.steps + args.warmup)
    host_rays = [tuple(t.pin_memory() for t in get_rays(poses[k], intr, H * world, W, rows=rows)) for k in range(n_res)]
    dev_rays = [(o.to(dev), d.to(dev)) for o, d in host_rays]
    kw = {}
    if wl == "mask":
        kw = dict(return_mask=1)

    def render_frame(ro, rd):
        if wl == "sam":
            # staged + return_feats is impossible in the reference API (`samvit.view(H, W, -1)` on a chunk),
            # so the feature frame is ONE non-staged call over all H*W rays with H, W of this rank's block -- the reference's own
            # call shape (trainer.py:536-537), which it can only afford at 64x64 because it materialises [N,32,128] tensors
            out = model.render(ro, rd, staged=False, perturb=False, return_feats=1, H=H, W=W, image_width=TILE_HINT and W)
            out = {k: (out[k].reshape(-1, out[k].shape[-1]) if k == "samvit" else out[k]) for k in keys}
        else:
            out = model.render(ro, rd, staged=True, perturb=False, image_width=TILE_HINT and W, **kw)
        if world > 1:
            out = gather_dict({k: out[k] for k in keys}, [n_local] * world)   # one NCCL all-gather of the packed outputs
        return out

    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_loop(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        n0 = _lib.launch_counter["n"]
        for i in range(steps):
            if flush is not None:
                flush.fill_(i & 0xFF)          # evict L2 between timed iterations (outside the per-step events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(warmup + i)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_counter["n"] - n0

    with torch.no_grad():
        # ---- device-resident arm ("value") ------------------------------------------------------
        sink = {}

        def step_dev(i):
            ro, rd = dev_rays[i % n_res]
            sink["out"] = render_frame(ro, rd)

        sampler = ClockSampler(local_rank)
        sampler.start()
        total_ms, launches = timed_loop(step_dev, args.steps, args.warmup)
        sampler.stop_flag = True
        ms_per_step = total_ms / args.steps
        value = n_total / (ms_per_step * 1e-3) / 1e6

        # ---- kernel-only timing of the dominant kernel (the fused render launch) for the roofline ----
        kern_ms = None
        if wl == "rgb":
            ro, rd = dev_rays[0]
            model.render(ro, rd, staged=True, perturb=False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(3, min(10, args.steps))
            tot = 0.0
            for i in range(reps):
                ro, rd = dev_rays[i % n_res]
                if flush is not None:
                    flush.fill_(i)
                a.record()
                model.render(ro, rd, staged=True, perturb=False)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            kern_ms = tot / reps

        # ---- end-to-end arm: host rays -> H2D -> render (public API) -> D2H of the image -----------
        # rank 0 reads the whole (gathered) frame back, like the reference's rank-0 image writer; the other ranks read nothing
        img_host = torch.empty(n_total, 3).pin_memory() if rank == 0 else None
        o_dev, d_dev = torch.empty(n_local, 3, device=dev), torch.empty(n_local, 3, device=dev)

        def step_e2e(i):
            ho, hd = host_rays[i % n_res]
            o_dev.copy_(ho, non_blocking=True)
            d_dev.copy_(hd, non_blocking=True)
            out = render_frame(o_dev, d_dev)
            if img_host is not None:
                img_host.copy_(out["image"], non_blocking=True)

        e2e_ms = timed_loop(step_e2e, args.steps, args.warmup)[0] / args.steps
        e2e_value = n_total / (e2e_ms * 1e-3) / 1e6

        # ---- informational: the same frame through render_image: host pose in (64 B), 8-bit image out ---------
        cam_value = None
        if wl == "rgb" and world == 1:
            u8_host = torch.empty(n_local, 3, dtype=torch.uint8).pin_memory()

            def step_cam(i):
                out = model.render_image(poses[i % N_POSES], intr, H, W, return_uint8=True)
                u8_host.copy_(out["image_u8"], non_blocking=True)

            cam_ms = timed_loop(step_cam, args.steps, args.warmup)[0] / args.steps
            cam_value = n_local / (cam_ms * 1e-3) / 1e6

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        k_ms = kern_ms if kern_ms is not None else ms_per_step
        achieved = BYTES_PER_RAY[wl] * n_local / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get(wl)   # from the committed ncu capture
        except Exception:
            pass
        line = {
            "metric": "Mrays/sec (RGB+feat render, 800x800)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl} {H}x{W} per GPU (global frame {H * world}x{W}), hashgrid L=16 T=2^19, MLP 2x64, "
                                   f"samples 128+64+32, one fused render launch per frame (reference: 4096 rays/batch)",
                       "rays_per_step": n_total, "poses": n_res, "parallelism": f"ray-row sharding x{world} + 1 all-gather",
                       "l2": "no flush" if flush is None else "L2 flushed between timed steps (256 MiB fill, outside the step events)",
                       "weights": "random init: seed-0 constructor, hash tables U(-1,1)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "sanerf::render_kernel" if wl == "rgb" else
                                   ("sanerf::render_kernel + sanerf::samvit_mlp_kernel" if wl == "sam" else
                                    "sanerf::render_kernel + sanerf::mask_head_kernel"),
                         "kernel_ms": k_ms, "algorithmic_bytes_per_ray": BYTES_PER_RAY[wl],
                         "mlp_tflops": FLOPS_PER_RAY[wl] * n_local / (k_ms * 1e-3) / 1e12},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * n_local * 12 * world,
                    "d2h_bytes_per_step": int(img_host.numel() * 4), "d2h": "rank 0 reads the gathered image",
                    "render_image_uint8_value": cam_value},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            mr, cores, sample, _ = cpu_baseline(wl, model.state_dict(), 40 if wl == "rgb" else 16, H, W)
            line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
