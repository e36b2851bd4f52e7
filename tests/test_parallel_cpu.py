"""CPU, gloo, world_size 2 (and 3, ragged): the multi-GPU plumbing of sanerf_hq_b200/parallel.py.

Rays are independent, so sharding must not change any per-ray value: every rank renders its contiguous block with a
deterministic stand-in render function (the real one needs a GPU; the GPU suite checks it) and after the single
all-gather every rank must hold exactly the single-process result, in the original ray order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sanerf_hq_b200.parallel import FrameGather, gather_dict, gather_rows, packed_floats_per_rank, render_sharded, shard_bounds


def _fake_render(rays_o, rays_d, **kw):
    """Per-ray function of the ray only (like the renderer): image [n,3], depth [n], weights_sum [n]."""
    # only IEEE-exact elementwise ops (no transcendental whose SIMD / scalar-tail code paths could differ by an ulp between a
    # slice and the full tensor)
    t = rays_o[:, 0] * rays_d[:, 0] + rays_o[:, 1] * rays_d[:, 1] + rays_o[:, 2] * rays_d[:, 2]
    return {"image": torch.stack([t * 2, t + 1, t * 0.5], dim=-1), "depth": t.abs(), "weights_sum": t * t, "num_points": 7}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        rays_o, rays_d = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
        full = _fake_render(rays_o, rays_d)
        got = render_sharded(_fake_render, rays_o, rays_d)
        ok = all(torch.equal(got[k], full[k]) for k in ("image", "depth", "weights_sum")) and set(got) == {"image", "depth", "weights_sum"}
        # gather_rows with explicit ragged counts, 2-D and 1-D payloads
        lo, hi = shard_bounds(n, world, rank)
        counts = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
        ok &= torch.equal(gather_rows(full["image"][lo:hi], counts), full["image"])
        ok &= torch.equal(gather_rows(full["depth"][lo:hi], counts), full["depth"])
        packed = gather_dict({k: full[k][lo:hi] for k in ("image", "depth", "weights_sum")}, counts)   # one collective
        ok &= all(torch.equal(packed[k], full[k]) for k in ("image", "depth", "weights_sum"))
        # a wide key (the 256-d SAM feature) next to the narrow ones: gathered on its own, not packed
        wide = torch.arange(n * 20, dtype=torch.float32).reshape(n, 20)
        mixed = gather_dict({"image": full["image"][lo:hi], "samvit": wide[lo:hi]}, counts)
        ok &= torch.equal(mixed["samvit"], wide) and torch.equal(mixed["image"], full["image"])
        # FrameGather (equal counts only): results land in the rank's slot of the full-frame tensors, in-place all-gather,
        # double-buffered frames, row groups
        if n % world == 0:
            n_local = n // world

            class FakeModel:
                def render(self, ro, rd, staged=True, out=None, **kw):
                    res = _fake_render(ro, rd)
                    for k, dst in out.items():
                        dst.copy_(res[k].reshape(dst.shape) if k in res else (ro[:, :1] * torch.arange(4.0)).reshape(dst.shape))
                    return res

            fg = FrameGather(n_local, {"image": (3,), "depth": (), "weights_sum": (), "wide": (4,)}, "cpu")
            assert fg.transport == "nccl"
            for frame, groups in enumerate((1, 2, 1)):
                scale = float(frame + 1)
                got = fg.render(FakeModel(), rays_o[lo:hi] * scale, rays_d[lo:hi], groups=groups)
                want = _fake_render(rays_o * scale, rays_d)
                ok &= all(torch.equal(got[k], want[k]) for k in ("image", "depth", "weights_sum"))
                ok &= torch.equal(got["wide"], rays_o[:, :1] * scale * torch.arange(4.0))
                if frame:
                    ok &= prev["image"].data_ptr() != got["image"].data_ptr()     # double-buffered
                prev = got
        torch.save(bool(ok), os.path.join(out_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 4096), (2, 4097), (3, 1000)])
def test_sharded_render_equals_single_process(tmp_path, world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(os.path.join(str(tmp_path), f"ok{r}.pt")) for r in range(world))


def test_packed_buffer_holds_only_the_narrow_keys():
    """the packed all-gather buffer must not reserve room for the wide keys that travel on their own."""
    widths = {"image": 3, "depth": 1, "weights_sum": 1, "samvit": 256, "instance_mask_logits": 2}
    assert packed_floats_per_rank(widths, [100, 99]) == 100 * 7


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 640000, 2560001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
