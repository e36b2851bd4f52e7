"""Ray-sharded multi-GPU rendering: one process per GPU, model replicated read-only, rays split into
contiguous row blocks, ONE all-gather of the composited outputs.

No collective touches the data path before compositing: rays are independent (every reduction in
`run` is along the sample axis).  The only exchange is the final gather of the per-ray results so
that every rank holds the whole frame, which is what the reference's (dead) eval-time
`dist.all_gather(preds)` does (nerf/trainer.py:1582-1585).
"""
import math

import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced [lo, hi) block of `n_items` for `rank`; the first `n_items % world_size` ranks get
    one extra item.  Works for ragged counts (n_items not divisible by world_size)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local, counts, group=None):
    """All-gather per-rank row blocks `local` [counts[rank], ...] into one [sum(counts), ...] tensor on every
    rank.  Equal counts -> a single all_gather_into_tensor (one NCCL collective over NVLink/NVSwitch); ragged
    counts -> padded to the largest block (still one collective) and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    tail = tuple(local.shape[1:])
    cmax = max(counts)
    if all(c == cmax for c in counts):
        out = local.new_empty((cmax * world,) + tail)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((cmax,) + tail)
    padded[: local.shape[0]] = local
    out = local.new_empty((cmax * world,) + tail)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * cmax: r * cmax + counts[r]] for r in range(world)], dim=0)


def gather_dict(parts, counts, group=None):
    """All-gather a dict of per-rank row blocks {key: [counts[rank], ...]} with ONE collective: the blocks are packed into one
    flat fp32 buffer per rank ([image | depth | weights_sum | ...]), gathered once over NVLink, and unpacked into full-frame
    tensors (rank-major row order, like `gather_rows`).  All tensors must be fp32; ragged counts are padded like `gather_rows`."""
    world = dist.get_world_size(group)
    if world == 1:
        return dict(parts)
    widths = {k: math.prod(parts[k].shape[1:]) for k in parts}   # floats per row
    # wide tensors (the 256-d SAM features: 1 KB per ray) go out on their own, straight from the tensor the kernel wrote --
    # packing them would cost two extra passes over hundreds of MB; the narrow per-ray outputs share one collective
    wide = {k: gather_rows(parts[k], counts, group) for k in parts if widths[k] > 16}
    keys = sorted(k for k in parts if k not in wide)
    if not keys:
        return wide
    cmax = max(counts)
    per_rank = cmax * sum(widths.values())
    ref = parts[keys[0]]
    send = ref.new_zeros(per_rank)
    off = 0
    for k in keys:
        n = parts[k].shape[0] * widths[k]
        send[off:off + n] = parts[k].reshape(-1)
        off += cmax * widths[k]
    recv = ref.new_empty(per_rank * world)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, per_rank)
    out, off = {}, 0
    for k in keys:
        w = widths[k]
        block = recv[:, off:off + cmax * w].reshape(world, cmax, *parts[k].shape[1:])
        out[k] = torch.cat([block[r, :counts[r]] for r in range(world)], dim=0) if any(c != cmax for c in counts) else \
            block.reshape(world * cmax, *parts[k].shape[1:])
        off += cmax * w
    out.update(wide)
    return out


def render_sharded(render_fn, rays_o, rays_d, group=None, keys=("image", "depth", "weights_sum"), **kwargs):
    """Render all rays [N,3] cooperatively.  Every rank passes the same full ray set (or at least its own block
    in the right place); rank r renders rows shard_bounds(N, world, r) with `render_fn(rays_o, rays_d, **kwargs)
    -> dict` (normally `model.render` with staged=True) and the listed outputs are all-gathered.
    The result is bit-identical to a single-process render because sharding does not change per-ray arithmetic."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = rays_o.shape[0]
    lo, hi = shard_bounds(N, world, rank)
    part = render_fn(rays_o[lo:hi], rays_d[lo:hi], **kwargs)
    if world == 1:
        return {k: part[k] for k in keys if k in part}
    counts = [shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world)]
    return gather_dict({k: part[k] for k in keys if k in part}, counts, group)   # ONE all-gather for all outputs
