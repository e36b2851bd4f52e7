"""Ray-sharded multi-GPU rendering: one process per GPU, model replicated read-only, rays split into contiguous row blocks,
the composited frame left on every rank.

No collective touches the data path before compositing: rays are independent (every reduction in `run` is along the sample
axis).  The only exchange is that of the per-ray results, the role of the reference's (dead) eval-time
`dist.all_gather(preds)` (nerf/trainer.py:1582-1601).  Two transports behind `FrameGather`:

* "peer"  (CUDA, one node) -- every rank owns full-frame buffers in NVLink peer memory (csrc/peer.cu, CUDA IPC).  A rank's
          rows reach the peers by copy-engine pushes (cudaMemcpyAsync peer copies on side streams), row group by row group
          while the SMs already render the next group -- what makes the 256-d SAM feature (1 KB per ray) affordable; a frame
          ends with one flag barrier.  No collective kernel, no pack / unpack copies, no SM spent on communication.
          Optionally (`kernel_stores=True`) the fused render kernel stores image / depth / weights_sum of its rays straight
          into every peer's buffer next to its own (`sanerf_render_args_t::peer_*`), i.e. the kernel's final stores are the
          all-gather of the narrow outputs; measured slightly slower than the pushes (see __init__), so off by default.
* "nccl"  (any backend, also gloo on CPU) -- the kernels store into the rank's slot of the full-frame tensors and one in-place
          `all_gather_into_tensor` per key gathers them (no pack / unpack copies either).

`gather_rows` / `gather_dict` remain for ragged row counts (padded all-gather).
"""
import ctypes
import math
import os

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


def packed_floats_per_rank(widths, counts, wide_threshold=16):
    """Size (floats) of one rank's block of `gather_dict`'s packed buffer: only the NARROW keys are packed."""
    return max(counts) * sum(w for w in widths.values() if w <= wide_threshold)


def gather_dict(parts, counts, group=None):
    """All-gather a dict of per-rank row blocks {key: [counts[rank], ...]} for RAGGED counts: the narrow blocks are packed into
    one flat fp32 buffer per rank ([image | depth | weights_sum | ...], padded to the largest block), gathered with one
    collective and unpacked (rank-major row order, like `gather_rows`).  Wide tensors (the 256-d SAM features: 1 KB per ray) go
    out on their own, straight from the tensor the kernel wrote.  All tensors must be fp32.  Equal counts: use FrameGather,
    which needs no pack / unpack copies at all."""
    world = dist.get_world_size(group)
    if world == 1:
        return dict(parts)
    widths = {k: math.prod(parts[k].shape[1:]) for k in parts}   # floats per row
    wide = {k: gather_rows(parts[k], counts, group) for k in parts if widths[k] > 16}
    keys = sorted(k for k in parts if k not in wide)
    if not keys:
        return wide
    cmax = max(counts)
    per_rank = packed_floats_per_rank(widths, counts)
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


NARROW_KEYS = ("image", "depth", "weights_sum")      # stored into the peers by the render kernel itself


class _DevMem:
    """Raw device memory as a __cuda_array_interface__ object (torch.as_tensor wraps it without copying or owning it)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class FrameGather:
    """Full-frame result buffers of a row-sharded render with EQUAL row counts per rank.

        fg = FrameGather(n_local, {"image": (3,), "depth": (), "weights_sum": (), "samvit": (256,)}, device)
        out = fg.render(model, rays_o_local, rays_d_local, return_feats=1, H=h, W=w)      # dict of FULL-frame tensors

    `render` makes the model's kernels store this rank's rows straight into its slot of the full-frame tensors (`out=`), fans
    them out (transport "peer": in-kernel peer stores + copy-engine pushes + one flag barrier; transport "nccl": in-place
    all-gathers) and returns the full-frame tensors, rank-major row order -- bit-identical to a single-process render of all
    rows because sharding does not change per-ray arithmetic.  Buffers are double-buffered: the tensors returned for frame i
    stay valid until frame i+2 is rendered, provided their consumers were enqueued on the current stream."""

    def __init__(self, n_local, spec, device, group=None, transport="auto", n_buffers=2, timeout_s=20.0, kernel_stores=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local, self.spec, self.device = int(n_local), {k: tuple(v) for k, v in spec.items()}, torch.device(device)
        self.n_buffers, self.frame, self.epoch, self.timeout_s = n_buffers, 0, 0, float(timeout_s)
        if transport == "auto":
            # peer memory pays where the payload is wide (the 256-d SAM feature: pushes overlap the rendering); for the narrow
            # per-ray outputs alone (20-28 B per ray) one in-place all-gather per key costs the same -- measured at 8 GPUs, rgb:
            # 11.47 ms per step (NCCL) vs 11.49 (copy-engine pushes) vs 11.67-11.72 (in-kernel peer stores)
            wide = any(math.prod(s) > 16 for s in self.spec.values())
            transport = os.environ.get("SANERF_TRANSPORT", "peer" if (wide and self.device.type == "cuda" and 1 < self.world <= 8) else "nccl")
        self.transport = transport if self.world > 1 else "local"
        self._peer = None
        # kernel_stores (or SANERF_PEER_KERNEL_STORES=1): the render kernel stores image / depth / weights_sum straight into the
        # peers' buffers; default off = the copy engines push them like the wide keys.  Measured at 8 GPUs (rgb, ms per step):
        # in-kernel stores 11.67-11.72, copy-engine pushes 11.49; at 2 GPUs the in-kernel stores showed erratic steps in one
        # of three runs (remote stores hold LSU slots while the peer's own gathers saturate its L2), the pushes did not.
        self._kernel_stores = (os.environ.get("SANERF_PEER_KERNEL_STORES", "0") == "1") if kernel_stores is None else bool(kernel_stores)
        n_total = self.n_local * self.world
        if self.transport == "peer":
            try:
                self._setup_peer(n_total)
            except Exception as e:   # IPC not permitted in this container / no peer access: say so and use the collective
                import warnings
                warnings.warn(f"sanerf_hq_b200.parallel: NVLink peer-memory transport unavailable ({e}); using in-place NCCL all-gathers",
                              RuntimeWarning)
                self.transport = "nccl"
            # every rank must agree on the transport (this all-reduce is also the "every rank has mapped every buffer" barrier)
            flag = torch.tensor([1 if self.transport == "peer" else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0 or self.transport != "peer":
                self._teardown_peer()
                self.transport = "nccl"
        if self.transport != "peer":
            self.buffers = [{k: torch.empty((n_total,) + s, device=self.device) for k, s in self.spec.items()}
                            for _ in range(n_buffers if self.world > 1 else 1)]

    # ---- peer-memory transport ------------------------------------------------------------------------------------------
    def _setup_peer(self, n_total):
        from . import _lib
        lib = _lib.load()
        self._lib, self._L = _lib, lib
        # layout of one buffer set: [flags 256 B][key 0 full frame][key 1 full frame] ... each 256-byte aligned
        self._off, off = {}, 256
        for k, s in self.spec.items():
            self._off[k] = off
            off += -(-n_total * math.prod(s) * 4 // 256) * 256
        self._set_bytes = off
        total = off * self.n_buffers
        with torch.cuda.device(self.device):
            # every rank takes part in the handle exchange even if its own allocation / export failed, so that a failure on
            # one rank makes ALL ranks fall back instead of dead-locking the others in a collective
            base, err, payload = ctypes.c_void_p(), None, None
            try:
                _lib.check(lib.sanerf_peer_alloc(total, ctypes.byref(base)), "sanerf_peer_alloc")
                handle = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
                _lib.check(lib.sanerf_peer_export(base, handle), "sanerf_peer_export")
                payload = bytes(handle.raw)
            except Exception as e:
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, payload, group=self.group)
            bases = [None] * self.world
            if err is None and all(h is not None for h in handles):
                bases[self.rank] = int(base.value)
                try:
                    for r in range(self.world):
                        if r != self.rank:
                            p = ctypes.c_void_p()
                            _lib.check(lib.sanerf_peer_open(ctypes.create_string_buffer(handles[r], _lib.PEER_HANDLE_BYTES), ctypes.byref(p)),
                                       "sanerf_peer_open")
                            bases[r] = int(p.value)
                except Exception as e:
                    err = e
            elif err is None:
                err = RuntimeError("a peer could not export its frame buffer")
            # SANERF_PUSH_SPLIT pieces per peer copy, each on its own stream: more copy engines in flight per push
            self._split = max(1, int(os.environ.get("SANERF_PUSH_SPLIT", "1")))
            self._peer = {"base": bases, "own": int(base.value or 0),
                          "streams": [torch.cuda.Stream(device=self.device) for _ in range((self.world - 1) * self._split)],
                          "status": torch.zeros(1, dtype=torch.int32, device=self.device)}
            if err is not None:
                raise err
            raw = torch.as_tensor(_DevMem(base.value, total), device=self.device)
        self.buffers = []
        for b in range(self.n_buffers):
            d = {}
            for k, s in self.spec.items():
                lo = b * self._set_bytes + self._off[k]
                d[k] = raw[lo:lo + n_total * math.prod(s) * 4].view(torch.float32).view((n_total,) + s)
            self.buffers.append(d)
        self._others = [r for r in range(self.world) if r != self.rank]

    def _teardown_peer(self):
        if self._peer is None:
            return
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self._peer["base"]):
            if r != self.rank and p:
                self._L.sanerf_peer_close(ctypes.c_void_p(p))
        self.buffers = None
        if self._peer["own"]:
            self._L.sanerf_peer_free(ctypes.c_void_p(self._peer["own"]))
        self._peer = None

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: every rank calls it)."""
        if self.transport == "peer" and self._peer is not None:
            dist.barrier(group=self.group)
            self._teardown_peer()

    def _peer_ptr(self, r, key, row):
        b = self.frame % self.n_buffers
        return self._peer["base"][r] + b * self._set_bytes + self._off[key] + row * math.prod(self.spec[key]) * 4

    def _push(self, key, lo, hi):
        """Copy-engine push of rows [lo, hi) (rank-local numbering) of `key` into every peer's frame buffer, ordered after the
        work enqueued so far on the current stream, on the side streams (one per peer)."""
        _lib, lib = self._lib, self._L
        row0 = self.rank * self.n_local + lo
        nbytes = (hi - lo) * math.prod(self.spec[key]) * 4
        ev = torch.cuda.Event()
        ev.record()
        n = len(self._others)
        for s in self._peer["streams"]:
            s.wait_event(ev)
        piece = -(-nbytes // self._split // 256) * 256
        for j in range(self._split):
            off, size = j * piece, min(piece, nbytes - j * piece)
            if size <= 0:
                break
            dst = (ctypes.c_void_p * n)(*[self._peer_ptr(r, key, row0) + off for r in self._others])
            streams = (ctypes.c_void_p * n)(*[s.cuda_stream for s in self._peer["streams"][j * n:(j + 1) * n]])
            _lib.check(lib.sanerf_peer_push(dst, ctypes.c_void_p(self._peer_ptr(self.rank, key, row0) + off), size, n, streams), "sanerf_peer_push")
        self._pushed = True

    def _barrier(self):
        _lib, lib = self._lib, self._L
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_pushed", False):
            for s in self._peer["streams"]:
                cur.wait_stream(s)
            self._pushed = False
        self.epoch += 1
        flags = (ctypes.c_void_p * self.world)(*[self._peer["base"][r] + (self.frame % self.n_buffers) * self._set_bytes for r in range(self.world)])
        _lib.check(lib.sanerf_peer_barrier(flags, self.rank, self.world, self.epoch, self.timeout_s, _lib.ptr(self._peer["status"]),
                                           ctypes.c_void_p(cur.cuda_stream)), "sanerf_peer_barrier")

    def check(self):
        """Synchronise and raise if a peer missed a barrier (diagnostic; not needed on the hot path)."""
        if self.transport == "peer":
            torch.cuda.synchronize(self.device)
            if int(self._peer["status"].item()):
                raise RuntimeError("sanerf_hq_b200.parallel: a peer did not reach the frame barrier in time")

    # ---- rendering ------------------------------------------------------------------------------------------------------
    def slot(self, key, lo=0, hi=None):
        """This rank's rows [lo, hi) of the current frame's full tensor `key` (what the kernels store into)."""
        hi = self.n_local if hi is None else hi
        base = self.rank * self.n_local
        return self.buffers[self.frame % len(self.buffers)][key][base + lo:base + hi]

    def render(self, model, rays_o, rays_d, groups=1, **kw):
        """model.render / run of this rank's `n_local` rays with the results fanned out; returns the full-frame tensors.
        `groups` > 1 renders the rows in that many (nearly equal) groups so that the copy-engine push of one group's wide
        tensors overlaps the rendering of the next."""
        n = self.n_local
        assert rays_o.shape[0] == n, "FrameGather.render: every rank renders exactly n_local rays"
        feats = kw.get("return_feats", 0)
        keys = list(self.spec)
        # group boundaries at multiples of 128 rays (whole tensor-core tiles), and of 4 image rows when the rays are an image
        # block of known width (keeps the kernel's 4x4-pixel tile traversal)
        unit = 128
        width = kw.get("image_width")
        if width and int(width) % 4 == 0 and n % (4 * int(width)) == 0:
            unit = math.lcm(128, 4 * int(width))
        n_units = n // unit if n % unit == 0 else 0
        groups = max(1, min(int(groups), n_units)) if n_units else 1
        bounds = [n if g == groups else (g * n_units // groups) * unit for g in range(groups + 1)] if groups > 1 else [0, n]
        kw.pop("H", None), kw.pop("W", None)
        for g in range(groups):
            lo, hi = bounds[g], bounds[g + 1]
            out = {k: self.slot(k, lo, hi) for k in keys}
            call = dict(kw, out=out)
            if self.transport == "peer" and self._kernel_stores:
                call["peer_out"] = {k: [self._peer_ptr(r, k, self.rank * n + lo) for r in self._others] for k in NARROW_KEYS if k in self.spec}
            if feats:
                # the reference's `samvit.view(H, W, -1)` (renderer.py:371-372) only needs H*W == rays of the call
                call.update(H=1, W=hi - lo)
                model.render(rays_o[lo:hi], rays_d[lo:hi], staged=False, **call)
            else:
                model.render(rays_o[lo:hi], rays_d[lo:hi], staged=True, **call)
            if self.transport == "peer":
                for k in keys:
                    if k not in NARROW_KEYS or not self._kernel_stores:
                        self._push(k, lo, hi)
        full = self.buffers[self.frame % len(self.buffers)]
        if self.transport == "peer":
            self._barrier()
        elif self.transport == "nccl":
            for k in keys:      # in place: the input is this rank's slot of the output
                dist.all_gather_into_tensor(full[k], self.slot(k), group=self.group)
        self.frame += 1
        return dict(full)       # "samvit" comes back flat, [world * n_local, 256]: the caller knows the frame's H, W


def render_sharded(render_fn, rays_o, rays_d, group=None, keys=("image", "depth", "weights_sum"), **kwargs):
    """Render all rays [N,3] cooperatively.  Every rank passes the same full ray set (or at least its own block
    in the right place); rank r renders rows shard_bounds(N, world, r) with `render_fn(rays_o, rays_d, **kwargs)
    -> dict` (normally `model.render` with staged=True) and the listed outputs are all-gathered.
    The result is bit-identical to a single-process render because sharding does not change per-ray arithmetic.
    (Ragged-capable convenience path; for per-frame use with equal row counts prefer FrameGather.)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = rays_o.shape[0]
    lo, hi = shard_bounds(N, world, rank)
    part = render_fn(rays_o[lo:hi], rays_d[lo:hi], **kwargs)
    if world == 1:
        return {k: part[k] for k in keys if k in part}
    counts = [shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world)]
    return gather_dict({k: part[k] for k in keys if k in part}, counts, group)   # ONE all-gather for all narrow outputs
