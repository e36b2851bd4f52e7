"""GPU: parallel.FrameGather on real devices.  One GPU: the "local" transport (results land in preallocated full-frame tensors,
bit-equal to a plain render).  Two or more GPUs (skipped on a single-GPU box): tests/multigpu_worker.py under torchrun -- the
NVLink peer-memory transport and the in-place NCCL transport must both reproduce the single-GPU frame bit for bit."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_frame_gather_single_gpu_writes_in_place():
    import bench
    from sanerf_hq_b200.parallel import FrameGather
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    H, W = 32, 128
    ro, rd = get_rays(orbit_pose(4).to(DEV), lego_intrinsics(H, W), H, W, device=DEV)
    with torch.no_grad():
        for wl, spec, kw in (("rgb", {}, {}), ("mask", {"instance_mask_logits": (2,)}, dict(return_mask=1)),
                             ("sam", {"samvit": (256,)}, dict(return_feats=1))):
            model = bench.build_model(wl, DEV)
            spec = dict({"image": (3,), "depth": (), "weights_sum": ()}, **spec)
            fg = FrameGather(H * W, spec, DEV)
            assert fg.transport == "local"
            for groups in (1, 4):
                got = fg.render(model, ro, rd, groups=groups, perturb=False, **kw)
                want = model.render(ro, rd, staged=wl != "sam", perturb=False, **(dict(kw, H=H, W=W) if wl == "sam" else kw))
                for k in spec:
                    assert got[k].data_ptr() == fg.buffers[0][k].data_ptr()
                    assert torch.equal(got[k], want[k].reshape(got[k].shape)), (wl, k, groups)
            del model


def test_frame_gather_multi_gpu_bit_exact():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tests", "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-6000:])
    sys.stderr.write(res.stderr[-3000:])
    assert res.returncode == 0, "multi-GPU frame differs from the single-GPU frame (see the JSON lines above)"
