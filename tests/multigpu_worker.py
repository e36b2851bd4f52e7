"""Worker of tests/test_parallel_gpu.py, launched with torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multigpu_worker.py

Every rank renders its row block of a frame through parallel.FrameGather -- transport "peer" (in-kernel NVLink stores + copy-
engine pushes + flag barrier) and transport "nccl" (in-place all-gathers) -- and compares the gathered frame BIT FOR BIT with
its own single-GPU render of the whole frame (sharding does not change per-ray arithmetic).  Prints one JSON line
per rank; exit code 0 only if everything matched."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import bench
    from sanerf_hq_b200.parallel import FrameGather
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    rows, W = 32, 128                      # rows per rank; 4096 rays per rank, 1024 per group of 4
    H = rows * world
    report = {"rank": rank, "world": world, "ok": True, "cases": {}}
    with torch.no_grad():
        for wl in ("rgb", "mask", "sam"):
            model = bench.build_model(wl, dev)
            spec = {"image": (3,), "depth": (), "weights_sum": ()}
            kw = {}
            if wl == "sam":
                spec["samvit"] = (256,)
                kw = dict(return_feats=1)
            if wl == "mask":
                spec["instance_mask_logits"] = (2,)
                kw = dict(return_mask=1)
            for transport, kstores in (("peer", False), ("peer", True), ("nccl", False)):   # CE pushes | in-kernel peer stores | NCCL
                fg = FrameGather(rows * W, spec, dev, transport=transport, kernel_stores=kstores)
                for frame, groups in enumerate((1, 4, 2)):
                    intr = lego_intrinsics(H, W)
                    ro, rd = get_rays(orbit_pose(3 + frame).to(dev), intr, H, W, device=dev)
                    lo, hi = rank * rows * W, (rank + 1) * rows * W
                    got = fg.render(model, ro[lo:hi].contiguous(), rd[lo:hi].contiguous(), groups=groups, perturb=False, **kw)
                    if wl == "sam":
                        want = model.render(ro, rd, staged=False, perturb=False, return_feats=1, H=H, W=W)
                        want["samvit"] = want["samvit"].reshape(-1, 256)
                    else:
                        want = model.render(ro, rd, staged=True, perturb=False, **kw)
                    fg.check()
                    same = {k: bool(torch.equal(got[k], want[k].reshape(got[k].shape))) for k in spec}
                    report["cases"][f"{wl}/{fg.transport}(asked {transport}, kernel_stores={kstores})/frame{frame}/groups{groups}"] = same
                    report["ok"] &= all(same.values())
                fg.close()
            del model
            torch.cuda.empty_cache()
    print(json.dumps(report), flush=True)
    ok = torch.tensor([1 if report["ok"] else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) else 1)


if __name__ == "__main__":
    main()
