"""Volumetric renderer (reference nerf/renderer.py:142-502) on top of libsanerf_b200.

`NeRFRenderer.render / run` keep the reference's signature, keyword arguments and result dict.
Two execution paths produce the same numbers:

* fused   -- eval / no-grad (every eval, test, decode and GUI call of nerf/trainer.py, with or without
             perturb): ONE persistent CUDA launch (`sanerf_render`, csrc/render.cu) does
             near/far, the three sampling stages with both proposal networks, sample_pdf,
             contraction, hash-grid lookups, the density / geometry MLP, SH, compositing and the
             deferred view MLP for all rays; `render(staged=True)` therefore does not chunk.
* composed -- anything that needs autograd (training): the same algorithm
             as differentiable torch ops around this package's CUDA encoders (forward+backward).

There is no CPU path: rays must be CUDA tensors.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib

try:  # training-only dependency of the reference (renderer.py:14)
    from torch_efficient_distloss import eff_distloss
except ImportError:  # same O(T) prefix-sum formulation (Sun et al., DVGOv2), differentiable torch ops
    def eff_distloss(w, m, interval):
        loss_uni = (1 / 3) * (interval * w.pow(2)).sum(dim=-1).mean()
        wm = w * m
        w_cum, wm_cum = w.cumsum(dim=-1), wm.cumsum(dim=-1)
        loss_bi = 2 * (wm[..., 1:] * w_cum[..., :-1] - w[..., 1:] * wm_cum[..., :-1]).sum(dim=-1).mean()
        return loss_bi + loss_uni


FUSED_NUM_STEPS = [128, 64, 32]


def distort_loss(bins, weights):
    """renderer.py:17-27"""
    with torch.amp.autocast("cuda", enabled=False):
        intervals = bins[..., 1:] - bins[..., :-1]
        mid = bins[..., :-1] + intervals / 2
        return eff_distloss(weights, mid, intervals)


def proposal_loss(all_bins, all_weights):
    """Inter-level (Mip-NeRF 360) bound loss of the proposal weights (renderer.py:30-57)."""
    with torch.amp.autocast("cuda", enabled=False):
        t_ref, w_ref = all_bins[-1].detach(), all_weights[-1].detach()
        total = 0
        for t, w in zip(all_bins[:-1], all_weights[:-1]):
            cw = torch.cat([torch.zeros_like(w[..., :1]), torch.cumsum(w, dim=-1)], dim=-1)
            last = w.shape[-1] - 1
            lo = (torch.searchsorted(t[..., :-1].contiguous(), t_ref[..., :-1].contiguous(), right=True) - 1).clamp(0, last)
            hi = torch.searchsorted(t[..., 1:].contiguous(), t_ref[..., 1:].contiguous(), right=True).clamp(0, last)
            bound = torch.take_along_dim(cw[..., 1:], hi, dim=-1) - torch.take_along_dim(cw[..., :-1], lo, dim=-1)
            total = total + ((w_ref - bound).clamp(min=0) ** 2 / (w_ref + 1e-8)).mean()
        return total


def contract(x):
    """L-inf scene contraction to (-2, 2) (renderer.py:60-69)."""
    with torch.amp.autocast("cuda", enabled=False):
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        mag, idx = x.abs().max(1, keepdim=True)
        scale = (1 / mag).repeat(1, shape[-1])
        scale.scatter_(1, idx, (2 - 1 / mag) / mag)
        return torch.where(mag < 1, x, x * scale).view(shape)


def uncontract(z):
    """Inverse of `contract` (renderer.py:72-81)."""
    with torch.amp.autocast("cuda", enabled=False):
        shape = z.shape
        z = z.reshape(-1, shape[-1])
        mag, idx = z.abs().max(1, keepdim=True)
        scale = 1 / (2 - mag.repeat(1, shape[-1])).clamp(min=1e-8)
        scale.scatter_(1, idx, 1 / (2 * mag - mag * mag).clamp(min=1e-8))
        return torch.where(mag < 1, z, z * scale).view(shape)


def sample_pdf(bins, weights, T, perturb=False):
    """Inverse-CDF resampling of `T` bin edges (renderer.py:84-119). bins [N,T0+1], weights [N,T0]."""
    with torch.amp.autocast("cuda", enabled=False):
        N, T0 = weights.shape
        weights = weights + 0.01
        pdf = weights / torch.sum(weights, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1).clamp(max=1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        u = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T).to(weights.device).expand(N, T)
        if perturb:
            u = u + (torch.rand_like(u) - 0.5) / T
        u = u.contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        below, above = torch.clamp(inds - 1, 0, T0), torch.clamp(inds, 0, T0)
        c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
        b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
        t = torch.clamp(torch.nan_to_num((u - c0) / (c1 - c0)), 0, 1)
        return b0 + t * (b1 - b0)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.05):
    """Slab test against aabb[6] (renderer.py:122-139) -> near [N,1], far [N,1]; miss -> 1e9."""
    with torch.amp.autocast("cuda", enabled=False):
        tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
        tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
        near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
        far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
        miss = far < near
        near = near.masked_fill(miss, 1e9)
        far = far.masked_fill(miss, 1e9)
        return torch.clamp(near, min=min_near), far


def _spacing(x):
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def _spacing_inv(x):
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.real_bound = opt.bound                       # world-space half extent for ray marching
        self.bound = 2 if self.opt.contract else opt.bound  # grid query range
        self.cascade = 1 + math.ceil(math.log2(self.bound))
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        box = torch.FloatTensor([-self.real_bound] * 3 + [self.real_bound] * 3)
        self.register_buffer("aabb_train", box)
        self.register_buffer("aabb_infer", box.clone())
        self._u_tables = {}
        self._aabb_host = (None, None)
        # set False to force the composed (torch op by op) path everywhere -- used by parity tests
        self.fused = True

    def forward(self, x, d, **kwargs):
        raise NotImplementedError()

    def density(self, x, **kwargs):
        raise NotImplementedError()

    def update_aabb(self, aabb):
        if not torch.is_tensor(aabb):
            aabb = torch.from_numpy(aabb).float()
        self.aabb_train = aabb.clamp(-self.real_bound, self.real_bound).to(self.aabb_train.device)
        self.aabb_infer = self.aabb_train.clone()
        self._aabb_host = (None, None)     # fresh tensors may reuse a freed block: never trust (data_ptr, version) across an update
        print(f"[INFO] update_aabb: {self.aabb_train.cpu().numpy().tolist()}")

    # ------------------------------------------------------------------------------------------
    # public entry points
    # ------------------------------------------------------------------------------------------
    def render(self, rays_o, rays_d, staged=False, cam_near_far=None, **kwargs):
        """rays_o, rays_d [N,3] -> dict(image [N,3], depth [N], weights_sum [N], ...) (renderer.py:185-219)."""
        if not staged:
            # reference quirk kept for identical results: `cam_near_far` is swallowed by this signature and NOT
            # forwarded in the non-staged call (renderer.py:187-188)
            return self.run(rays_o, rays_d, **kwargs)
        if self._can_fuse(rays_o, kwargs) and not kwargs.get("return_feats", 0):
            # one persistent launch over all rays; max_ray_batch chunking only bounds the temporaries of the
            # object head
            return self._run_fused(rays_o, rays_d, cam_near_far=cam_near_far, noise_chunk=self.opt.max_ray_batch, **kwargs)
        N, device = rays_o.shape[0], rays_o.device
        results = {}
        out = kwargs.pop("out", None)
        step = self.opt.max_ray_batch
        for head in range(0, N, step):
            tail = min(head + step, N)
            cnf = cam_near_far
            if cnf is not None and cnf.shape[0] != 1:
                cnf = cnf[head:tail]
            part = self.run(rays_o[head:tail], rays_d[head:tail], cam_near_far=cnf, **kwargs)
            for k, v in part.items():
                if v is None:
                    continue
                if torch.is_tensor(v):
                    if k not in results:
                        results[k] = torch.empty(N, *v.shape[1:], device=device)
                    results[k][head:tail] = v
                else:
                    results[k] = v
        if out:
            for k, dst in out.items():
                if k in results:
                    dst.copy_(results[k].reshape(dst.shape))
                    results[k] = dst
        return results

    @torch.no_grad()
    def render_image(self, pose, intrinsics, H, W, rows=None, return_uint8=False, **kwargs):
        """Render image rows [rows[0], rows[1]) (default: the whole H x W frame) of a pinhole camera WITHOUT materialising rays:
        the fused kernel derives every ray from `pose` (cam2world [4,4] or [3,4]) and `intrinsics` (fx, fy, cx, cy) exactly
        like the full-image branch of the reference's nerf/utils.py::get_rays (:262-287).  Same kwargs / result dict as
        `render`; `return_uint8=True` adds `image_u8` [N,3] = (image * 255) cast like trainer.py:1140-1143.
        Eval / no-grad only (the fused path); the model's device is used."""
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("NeRFRenderer.render_image: the model must be on a CUDA device (there is no CPU path)")
        kw = dict(kwargs)
        kw.setdefault("perturb", False)
        if not self._can_fuse(torch.empty(0, device=device), kw):
            raise RuntimeError("NeRFRenderer.render_image needs the fused path (eval / no-grad, default sample counts)")
        r0, r1 = (0, H) if rows is None else rows
        pose = torch.as_tensor(pose, dtype=torch.float32).reshape(-1, 4)[:3].cpu()
        camera = (pose.reshape(-1).tolist(), [float(v) for v in intrinsics], int(W), int(r0) * int(W), (int(r1) - int(r0)) * int(W), device)
        kw.pop("staged", None)
        if kw.get("return_feats", 0):
            kw.setdefault("H", r1 - r0)
            kw.setdefault("W", W)
        return self._run_fused(None, None, camera=camera, return_uint8=return_uint8, **kw)

    def run(self, rays_o, rays_d, bg_color=None, perturb=False, cam_near_far=None, update_proposal=True,
            return_feats=0, return_mask=0, H=None, W=None, image_width=None, out=None, peer_out=None, max_ctas=0,
            feature_layout=None, feature_size=None, **kwargs):
        """Not in the reference: `image_width` -- optional hint that the rays are a row-major image block of that width, which
        lets the fused kernel walk 4x4-pixel tiles (results are unchanged); `out` -- optional dict of preallocated result
        tensors (`image`, `depth`, `weights_sum`, `samvit`, `instance_mask_logits`) the kernels store into directly (used by
        parallel.FrameGather so that a rank's pixels land in its slot of the all-gather buffer without a copy); `peer_out` -- dict
        key -> list of raw device pointers (the other ranks' frame buffers in NVLink peer memory, at ray 0 of this call) that
        the fused kernel additionally stores image / depth / weights_sum into; `max_ctas` -- cap on the persistent CTAs;
        `feature_layout="nchw"` (+ optional `feature_size=(Ho, Wo)`) -- return the SAM feature frame as `samvit_nchw`
        [1,256,Ho,Wo] instead of `samvit` [H,W,256]: the reshape / permute / contiguous / F.interpolate(bilinear) chain of the
        reference's consumer (nerf/trainer.py:540-546) done by the head's epilogue (+ one fused resize pass when the size
        changes) -- SURVEY.md 8f-3."""
        if self.opt.render_mesh:
            return {}  # the reference's mesh branch is commented out and returns an empty dict (renderer.py:386-498)
        kw = dict(bg_color=bg_color, perturb=perturb, cam_near_far=cam_near_far, update_proposal=update_proposal,
                  return_feats=return_feats, return_mask=return_mask, H=H, W=W, image_width=image_width,
                  feature_layout=feature_layout, feature_size=feature_size)
        if self._can_fuse(rays_o, kw):
            return self._run_fused(rays_o, rays_d, out=out, peer_out=peer_out, max_ctas=max_ctas, **kw)
        if peer_out:
            raise RuntimeError("NeRFRenderer.run: peer_out needs the fused kernel (eval / no-grad)")
        if self._can_train_heads_on_fused_geometry(rays_o, kw):
            res = self._run_frozen_geometry(rays_o, rays_d, **kw)
            if out:
                for k, dst in out.items():
                    if k in res:
                        dst.copy_(res[k].detach().reshape(dst.shape))
            return res
        if not torch.is_grad_enabled() and not self.training:
            self._warn_composed(kw)
        res = self._run_composed(rays_o, rays_d, **kw)
        if out:
            for k, dst in out.items():
                if k in res:
                    dst.copy_(res[k].reshape(dst.shape))
                    res[k] = dst
        return res

    def _can_train_heads_on_fused_geometry(self, rays_o, kw):
        """Object / SAM stage training (nerf/trainer.py:401-409, 507-550) after main.py's name-based freezing (main.py:249-256): no
        parameter of the geometry / colour field requires grad, and the heads' losses cannot reach it anyway (`weights.detach()`,
        `geo_feat.detach()`, renderer.py:378-384).  The geometry then comes from the fused kernel (no-grad) and only the heads --
        feature grid + MLP -- run as differentiable ops on the samples it placed."""
        if not (torch.is_grad_enabled() and rays_o.is_cuda and not rays_o.requires_grad) or kw.get("perturb", False):
            return False
        feats = self.opt.with_sam and kw.get("return_feats", 0)
        if not (kw.get("return_mask", 0) or feats) or not (self.opt.with_mask or self.opt.with_sam):
            return False
        if self._fuse_blocker(kw) is not None or kw.get("feature_layout") not in (None, "nhwc"):
            return False
        frozen = [self.grid, self.grid_mlp, self.view_mlp, self.prop_encoders, self.prop_mlp]
        return not any(q.requires_grad for m in frozen for q in m.parameters())

    def _run_frozen_geometry(self, rays_o, rays_d, bg_color=None, cam_near_far=None, return_feats=0, return_mask=0, H=None, W=None,
                             **kwargs):
        from .encoders import grid_encode
        taps = {"weights2": None, "f_image": None}
        base = self._run_fused(rays_o, rays_d, bg_color=bg_color, cam_near_far=cam_near_far, records_only=True, taps=taps)
        N = rays_o.shape[0]
        rec, w = base.pop("_records"), base.pop("_weights")
        x01, geo_feat = rec[:, :3].contiguous(), rec[:, 3:].reshape(N, 32, 15)
        results = {k: base[k] for k in ("weights_sum", "depth", "image")}

        def features(enc):    # enc(xyzs, bound) on the points the kernel sampled, already mapped to [0,1]^3 (bound 0: no mapping)
            return grid_encode(x01, enc.embeddings, enc.offsets, enc.per_level_scale, enc.base_resolution, False, enc.gridtype_id,
                               enc.align_corners, enc.interp_id, None, 0.0).view(N, 32, -1)

        if self.opt.with_sam:
            f_sam = torch.sum(w.unsqueeze(-1) * features(self.s_grid), dim=-2)                                  # renderer.py:361
            f = torch.cat([f_sam, taps["f_image"], results["image"], results["depth"].unsqueeze(-1)], dim=-1)  # :363
            samvit = self.samvit_mlp(f)
            if return_feats > 0:
                results["samvit"] = samvit.view(H, W, -1)
        if return_mask > 0:
            point_masks = self.mask_mlp(torch.cat([features(self.m_grid), geo_feat], dim=-1))                   # :378-381
            results["instance_mask_logits"] = torch.sum(w.unsqueeze(-1) * point_masks, dim=-2)                 # :384
        return results

    def _warn_composed(self, kw):
        """Eval-mode call that cannot take the single-launch kernel: say so once per reason instead of silently running ~10x
        slower on the op-by-op path."""
        reason = self._fuse_blocker(kw)
        if reason is None:
            return
        seen = self.__dict__.setdefault("_composed_warned", set())
        if reason not in seen:
            seen.add(reason)
            import warnings
            warnings.warn(f"sanerf_hq_b200: eval-mode render falls back to the op-by-op torch path ({reason}); "
                          "the fused sm_100a kernel is not used for this call", RuntimeWarning, stacklevel=3)

    # ------------------------------------------------------------------------------------------
    # fused path
    # ------------------------------------------------------------------------------------------
    def _fuse_blocker(self, kw):
        """None when the configuration is one the fused kernel implements, else the reason it is not (a string)."""
        if not self.fused:
            return "model.fused is False"
        if kw.get("perturb", False) and self.opt.with_sam and kw.get("return_feats", 0):
            return "perturb=True together with the SAM feature head"
        if self.opt.render_mesh:
            return "render_mesh"
        if list(self.opt.num_steps) != FUSED_NUM_STEPS:
            return f"num_steps {list(self.opt.num_steps)} != {FUSED_NUM_STEPS}"
        if self.opt.with_sam and not self.opt.sam_use_view_direction:
            return "with_sam without sam_use_view_direction"
        if kw.get("return_mask", 0) and self.opt.mask_mlp_type != "default":
            return f"mask_mlp_type {self.opt.mask_mlp_type}"
        shape = (self.grid.num_levels, self.prop_encoders[0].num_levels, self.grid_mlp.dim_hidden, self.view_mlp.dim_hidden)
        if shape not in ((16, 5, 64, 32), (4, 4, 16, 16)) or self.prop_encoders[1].num_levels != self.prop_encoders[0].num_levels:
            return f"network shape {shape} has no fused instantiation"
        return None

    def _can_fuse(self, rays_o, kw):
        if not rays_o.is_cuda or torch.is_grad_enabled():
            return False
        if self.training and not self.opt.with_mask and not self.opt.with_sam:
            return False  # rgb training mode also returns losses / weights (renderer.py:343-351)
        return self._fuse_blocker(kw) is None

    def _u_table(self, T, device):
        key = (T, str(device))
        if key not in self._u_tables:  # the reference builds u on the CPU and copies it over (renderer.py:97)
            self._u_tables[key] = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T).to(device)
        return self._u_tables[key]

    @staticmethod
    def _fill_grid(dst, enc):
        dst.embeddings = enc.embeddings.data_ptr()
        dst.num_levels = enc.num_levels
        dst.level_dim = enc.level_dim
        for i, o in enumerate(enc.offsets_host):
            dst.offset[i] = o
        for i, r in enumerate(enc.level_resolutions()):
            dst.res[i] = r

    def _model_struct(self, device):
        """sanerf_model_t for the current parameters.  Filling the ctypes struct costs ~200 field stores; it is rebuilt only when
        a tensor it points to moved (load_state_dict / .to() / optimizer steps keep data_ptr), the mode or the box changed."""
        tensors = [self.grid.embeddings, self.prop_encoders[0].embeddings, self.prop_encoders[1].embeddings]
        tensors += [l.weight for net in (self.grid_mlp, self.view_mlp, self.prop_mlp[0], self.prop_mlp[1]) for l in net.net]
        if self.opt.with_sam:
            tensors.append(self.s_grid.embeddings)
        if self.opt.with_mask and self.opt.mask_mlp_type == "default":
            tensors.append(self.m_grid.embeddings)
        box = self.aabb_train if self.training else self.aabb_infer
        ckey = (tuple(t.data_ptr() for t in tensors), box.data_ptr(), box._version, str(device), self.min_near, float(self.bound),
                bool(self.opt.contract), self.opt.background, getattr(self.opt, "n_inst", 0), self._aabb_host[0])
        cached = self.__dict__.get("_model_cache")
        if cached is not None and cached[0] == ckey:
            return cached[1], list(cached[2])
        m = _lib.ModelT()
        keep = []
        for i in range(2):
            self._fill_grid(m.prop_grid[i], self.prop_encoders[i])
            m.prop_w0[i] = self.prop_mlp[i].net[0].weight.data_ptr()
            m.prop_w1[i] = self.prop_mlp[i].net[1].weight.data_ptr()
        self._fill_grid(m.grid, self.grid)
        for i in range(3):
            m.grid_w[i] = self.grid_mlp.net[i].weight.data_ptr()
            m.view_w[i] = self.view_mlp.net[i].weight.data_ptr()
        m.grid_hidden = self.grid_mlp.dim_hidden
        m.view_hidden = self.view_mlp.dim_hidden
        if self.opt.with_sam:
            self._fill_grid(m.s_grid, self.s_grid)
        if self.opt.with_mask and self.opt.mask_mlp_type == "default":
            self._fill_grid(m.m_grid, self.m_grid)
            m.n_inst = self.opt.n_inst
        key = (box.data_ptr(), box._version)
        if self._aabb_host[0] != key:  # host copy cached so a render call does not sync the stream
            self._aabb_host = (key, box.tolist())
        aabb = self._aabb_host[1]
        for i in range(6):
            m.aabb[i] = aabb[i]
        m.min_near = self.min_near
        m.grid_bound = float(self.bound)
        m.contract = int(bool(self.opt.contract))
        m.last_sample_opaque = int(self.opt.background == "last_sample")
        u65, u33 = self._u_table(65, device), self._u_table(33, device)
        m.u65, m.u33 = u65.data_ptr(), u33.data_ptr()
        keep += [u65, u33]
        self.__dict__["_model_cache"] = ((ckey[:-1] + (self._aabb_host[0],)), m, list(keep))
        return m, keep

    @torch.no_grad()
    def _run_fused(self, rays_o, rays_d, bg_color=None, perturb=False, cam_near_far=None, update_proposal=True,
                   return_feats=0, return_mask=0, H=None, W=None, taps=None, camera=None, return_uint8=False, image_width=None,
                   out=None, peer_out=None, max_ctas=0, feature_layout=None, feature_size=None, noise_chunk=None, records_only=False,
                   **kwargs):
        if camera is None:
            rays_o = rays_o.contiguous().float()
            rays_d = rays_d.contiguous().float()
            _lib.require_cuda(rays_o, rays_d, what="NeRFRenderer.run")
            N, device = rays_o.shape[0], rays_o.device
        else:   # rays are generated inside the kernel: camera = (pose 3x4 as 12 floats, (fx, fy, cx, cy), width, first linear pixel index, n_rays, device)
            N, device = camera[4], camera[5]
        lib = _lib.load()
        for q in self.parameters():
            if q.device != device:
                raise RuntimeError("NeRFRenderer.run: model and rays must be on the same CUDA device")
            break
        model, keep = self._model_struct(device)

        def result(key, *shape):
            """Result tensor `key`: the caller's preallocated one (`out`) or a fresh one."""
            t = out.get(key) if out else None
            if t is None:
                return torch.empty(*shape, device=device)
            if tuple(t.shape) != shape or t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
                raise RuntimeError(f"NeRFRenderer.run: out['{key}'] must be a contiguous fp32 tensor of shape {shape} on {device}")
            return t

        image = result("image", N, 3)
        depth = result("depth", N)
        weights_sum = result("weights_sum", N)
        results = {"weights_sum": weights_sum, "depth": depth, "image": image}

        a = _lib.RenderArgsT()
        a.N = N
        if camera is None:
            a.rays_o, a.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
        else:
            a.cam_w, a.cam_ray0 = int(camera[2]), int(camera[3])
            for i, v in enumerate(camera[1]):
                a.cam_intrinsics[i] = float(v)
            for i, v in enumerate(camera[0]):
                a.cam_pose[i] = float(v)
        if camera is not None:
            image_width = camera[2]
        # traversal hint (rays are a row-major image block of this width): only when no per-chunk pointer arithmetic is involved
        if image_width and not return_mask and N % (4 * int(image_width)) == 0 and int(image_width) % 4 == 0:
            a.tile_w = int(image_width)
        if return_uint8:
            image_u8 = torch.empty(N, 3, device=device, dtype=torch.uint8)
            results["image_u8"] = image_u8
            a.image_u8 = image_u8.data_ptr()
        if cam_near_far is not None:
            cnf = cam_near_far.to(device=device, dtype=torch.float32).contiguous()
            keep.append(cnf)
            a.cam_near_far, a.cam_near_far_rows = cnf.data_ptr(), cnf.shape[0]
        a.bg_scalar = 1.0
        if bg_color is not None:
            if torch.is_tensor(bg_color) and bg_color.numel() > 1:
                bg = bg_color.to(device=device, dtype=torch.float32).reshape(-1, 3).contiguous()
                keep.append(bg)
                a.bg_color, a.bg_rows = bg.data_ptr(), bg.shape[0]
            else:
                a.bg_scalar = float(bg_color)
        a.image, a.depth, a.weights_sum = image.data_ptr(), depth.data_ptr(), weights_sum.data_ptr()
        a.max_ctas = int(max_ctas or 0)
        if getattr(self, "_render_ws", None) is None or self._render_ws.device != device:
            self._render_ws = torch.empty(lib.sanerf_render_workspace_bytes(), dtype=torch.uint8, device=device)
        a.workspace = self._render_ws.data_ptr()
        peer = None
        if peer_out:
            peer = [peer_out["image"], peer_out["depth"], peer_out["weights_sum"]]
            a.n_peer_out = len(peer[0])

        def set_peer(head):
            """Peer pointers of ray `head` of this call (image 12 B, depth / weights_sum 4 B per ray)."""
            if peer:
                for i in range(a.n_peer_out):
                    a.peer_image[i] = peer[0][i] + 12 * head
                    a.peer_depth[i] = peer[1][i] + 4 * head
                    a.peer_weights_sum[i] = peer[2][i] + 4 * head

        set_peer(0)
        if perturb:
            # perturb=True (renderer.py:267-270, 99-100): the jitter is drawn HERE with torch.rand, in the reference's order -- per
            # chunk of `noise_chunk` rays (what one `run` call of the reference's staged loop sees): [n,129], then [n,65], [n,33] --
            # so that under the same torch seed the fused kernel consumes exactly the reference's random stream
            step = int(noise_chunk or N) or 1
            noise = [torch.empty(N, t, device=device) for t in (129, 65, 33)]
            for head in range(0, N, step):
                for t in noise:
                    t[head:head + step] = torch.rand(min(step, N - head), t.shape[1], device=device)
            a.noise0, a.noise1, a.noise2 = (t.data_ptr() for t in noise)
            keep += noise

        want_sam = self.opt.with_sam and return_feats > 0
        want_mask = return_mask > 0
        if records_only:      # geometry only: per-sample records (point, geo_feat) + weights for differentiable heads on top
            want_sam, want_mask = False, True
        if taps is not None:  # parity taps (tests): dict name -> None, filled with tensors
            shapes = {"inds0": ((N, 65), torch.int16), "inds1": ((N, 33), torch.int16), "weights2": ((N, 32), torch.float32),
                      "sigma2": ((N, 32), torch.float32), "bins2": ((N, 33), torch.float32), "f_image": ((N, 31), torch.float32)}
            for name in list(taps):
                shp, dt = shapes[name]
                # rows rounded up to whole 128-sample tiles (4 rays): the tensor-core object head reads full tiles of weights2
                taps[name] = torch.empty((-(-N // 4) * 4,) + shp[1:], device=device, dtype=dt)[:N]
                setattr(a, name, taps[name].data_ptr())

        if not want_mask:
            sam_in = None
            if want_sam:
                sam_in = self._alloc_rows_padded(N, self.samvit_mlp[0].dim_in, device)
                a.sam_in = sam_in.data_ptr()
            with torch.cuda.device(device), _lib.timed("render_kernel"):
                _lib.check(lib.sanerf_render(ctypes.byref(model), ctypes.byref(a), _lib.stream_ptr()), "sanerf_render")
                _lib.count_launch(2)     # weight prepare + the persistent render kernel
            if want_sam:
                results.update(self._feature_results(sam_in, H, W, out, feature_layout, feature_size))
            return results

        # object head.  Default shapes: the fused kernel leaves an 18-float record (point, geo_feat) per sample and the
        # tensor-core head (csrc/heads.cu) gathers m_grid, runs mask_mlp 143 -> 256 -> 256 -> n_inst and composites, one launch
        # per chunk of rays (the chunk bounds the scratch).  Other widths (config #1's small network, n_inst > 16): the fused
        # kernel writes the full per-sample inputs cat[m_grid(x), geo_feat] and the nn.Module MLP consumes them.
        if records_only:
            n_pad = -(-max(N, 1) // 4) * 4
            mask_in = torch.empty(n_pad * 32 * 18, device=device)
            w2 = torch.empty(n_pad, 32, device=device)
            a.mask_in, a.mask_in_tiled = mask_in.data_ptr(), 2
            if not a.weights2:
                a.weights2 = w2.data_ptr()
            else:
                w2 = taps["weights2"]
            with torch.cuda.device(device), _lib.timed("render_kernel"):
                _lib.check(lib.sanerf_render(ctypes.byref(model), ctypes.byref(a), _lib.stream_ptr()), "sanerf_render")
                _lib.count_launch(2)
            # records are tile-transposed [tile][18][128] (row = ray * 32 + sample): back to [N * 32, 18]
            results["_records"] = mask_in.view(-1, 18, 128).permute(0, 2, 1).reshape(-1, 18)[:N * 32]
            results["_weights"] = w2[:N]
            return results
        n_inst = self.opt.n_inst
        logits = result("instance_mask_logits", N, n_inst)
        width = self.mask_mlp[0].dim_in
        net = self.mask_mlp[0].net
        tc_head = (width == 143 and len(net) == 3 and net[0].weight.shape == (256, 143) and net[1].weight.shape == (256, 256)
                   and net[2].weight.shape == (n_inst, 256) and n_inst <= 16 and all(l.bias is None for l in net))
        tc_head = tc_head and self.m_grid.num_levels == 16 and self.m_grid.level_dim == 8
        if not tc_head:
            seen = self.__dict__.setdefault("_composed_warned", set())
            if "mask_head" not in seen and width == 143:       # default-size m_grid but a head the tensor-core kernel does not cover
                seen.add("mask_head")
                import warnings
                warnings.warn("sanerf_hq_b200: the object head runs as torch nn.Linear layers (the tensor-core head covers 143 -> 256 -> "
                              f"256 -> n_inst <= 16 without bias; this model: n_inst = {n_inst}, layers "
                              f"{[tuple(l.weight.shape) for l in net]})", RuntimeWarning, stacklevel=3)
        # rays per launch pair (bounds the scratch: 2.3 KB per ray of records); a multiple of 4 rays = whole 128-sample tiles,
        # so the chunking is invisible in the results.  `opt.mask_chunk_rays` overrides (tests).
        cap = int(getattr(self.opt, "mask_chunk_rays", 0) or (1 << 20 if tc_head else int(getattr(self.opt, "max_ray_batch", 4096)) * 8))
        chunk = max(4, min(-(-N // 4) * 4, cap // 4 * 4))
        mask_in = torch.empty(chunk * 32 * (18 if tc_head else width), device=device)
        w2 = torch.empty(chunk, 32, device=device)
        if tc_head:
            if getattr(self, "_mask_ws", None) is None or self._mask_ws.device != device:
                self._mask_ws = torch.empty(lib.sanerf_mask_head_workspace_bytes(), dtype=torch.uint8, device=device)
            mw = [l.weight.detach().contiguous() for l in net]
            a.mask_in_tiled = 2
        sam_full = self._alloc_rows_padded(N, self.samvit_mlp[0].dim_in, device) if want_sam else None
        base = {f: getattr(a, f) for f in ("rays_o", "rays_d", "image", "depth", "weights_sum", "cam_near_far", "bg_color",
                                           "inds0", "inds1", "weights2", "sigma2", "bins2", "f_image", "image_u8", "noise0", "noise1", "noise2")}
        strides = {"rays_o": 12, "rays_d": 12, "image": 12, "depth": 4, "weights_sum": 4, "inds0": 130, "inds1": 66,
                   "weights2": 128, "sigma2": 128, "bins2": 132, "f_image": 124, "image_u8": 3, "noise0": 516, "noise1": 260, "noise2": 132}
        cam_ray0 = a.cam_ray0
        user_w2 = base["weights2"]
        for head in range(0, N, chunk):
            n = min(chunk, N - head)
            for f, s in strides.items():
                if base[f]:
                    setattr(a, f, base[f] + head * s)
            if base["cam_near_far"] and a.cam_near_far_rows > 1:
                a.cam_near_far = base["cam_near_far"] + head * 8
            if base["bg_color"] and a.bg_rows > 1:
                a.bg_color = base["bg_color"] + head * 12
            a.N = n
            a.cam_ray0 = cam_ray0 + head
            set_peer(head)
            a.mask_in = mask_in.data_ptr()
            if not user_w2:
                a.weights2 = w2.data_ptr()
            if want_sam:
                a.sam_in = sam_full.data_ptr() + head * sam_full.shape[1] * 4
            wts = taps["weights2"][head:head + n] if user_w2 else w2[:n]
            with torch.cuda.device(device):
                with _lib.timed("render_kernel"):
                    _lib.check(lib.sanerf_render(ctypes.byref(model), ctypes.byref(a), _lib.stream_ptr()), "sanerf_render")
                    _lib.count_launch(2)
                if tc_head:
                    dst = logits[head:head + n]
                    with _lib.timed("mask_head_kernel"):
                        _lib.check(lib.sanerf_mask_head(_lib.ptr(mask_in), _lib.ptr(wts), ctypes.byref(model.m_grid), _lib.ptr(mw[0]),
                                                        _lib.ptr(mw[1]), _lib.ptr(mw[2]), n_inst, n, _lib.ptr(self._mask_ws), _lib.ptr(dst),
                                                        _lib.stream_ptr()), "sanerf_mask_head")
                        _lib.count_launch(2)
            if not tc_head:
                point_masks = self.mask_mlp(mask_in[:n * 32 * width].view(n, 32, width))
                logits[head:head + n] = torch.sum(wts.unsqueeze(-1) * point_masks, dim=-2)
        if want_sam:
            results.update(self._feature_results(sam_full, H, W, out, feature_layout, feature_size))
        results["instance_mask_logits"] = logits
        return results

    @staticmethod
    def _alloc_rows_padded(n, width, device, multiple=128):
        """[n, width] view of a buffer with whole tiles of `multiple` rows (the tensor-core heads read full tiles)."""
        return torch.empty(-(-max(n, 1) // multiple) * multiple, width, device=device)[:n]

    def _feature_results(self, f, H, W, out, layout, size):
        """The SAM feature frame from the composited head inputs f [H*W,163]: `samvit` [H,W,256] like the reference
        (renderer.py:371-374), or with layout "nchw" `samvit_nchw` [1,256,Ho,Wo] = permute + bilinear resize of it
        (trainer.py:540-546) without materialising the intermediate tensors."""
        if layout in (None, "nhwc"):
            return {"samvit": self._samvit_head(f, out=out.get("samvit") if out else None).view(H, W, -1)}
        if layout != "nchw":
            raise ValueError(f"feature_layout must be None, 'nhwc' or 'nchw', not {layout!r}")
        Ho, Wo = (int(v) for v in (size or (H, W)))
        if (Ho, Wo) == (H, W):          # bilinear resampling to the same size is the identity: the head stores channel-major
            return {"samvit_nchw": self._samvit_head(f, nchw=True).view(1, -1, H, W)}
        nhwc = self._samvit_head(f)
        if not nhwc.is_cuda:
            return {"samvit_nchw": torch.nn.functional.interpolate(nhwc.view(1, H, W, -1).permute(0, 3, 1, 2).contiguous(), (Ho, Wo),
                                                                   mode="bilinear")}
        C = nhwc.shape[-1]
        res = torch.empty(1, C, Ho, Wo, device=nhwc.device)
        with torch.cuda.device(nhwc.device), _lib.timed("feature_resize_nchw_kernel"):
            _lib.check(_lib.load().sanerf_feature_resize_nchw(_lib.ptr(nhwc), H, W, C, Ho, Wo, _lib.ptr(res), _lib.stream_ptr()),
                       "sanerf_feature_resize_nchw")
            _lib.count_launch()
        return {"samvit_nchw": res}

    def _samvit_head(self, f, out=None, nchw=False):
        """samvit_mlp (SkipConnMLP + LayerNorm, network.py:113-116) of the composited per-ray features f [n,163].
        No-grad CUDA input with the reference's shapes -> the tensor-core head (csrc/heads.cu); anything else -> nn.Modules.
        `out`: optional preallocated contiguous fp32 tensor with n*256 elements the head stores into."""
        mlp, ln = self.samvit_mlp[0], self.samvit_mlp[1]
        net = mlp.net
        ok = (f.is_cuda and not torch.is_grad_enabled() and f.dim() == 2 and f.shape[1] == 163 and len(net) == 5
              and list(mlp.skip_layers) == [2] and [tuple(l.weight.shape) for l in net] ==
              [(256, 163), (256, 256), (256, 419), (256, 256), (256, 256)] and all(l.bias is not None for l in net)
              and tuple(ln.normalized_shape) == (256,) and abs(ln.eps - 1e-5) < 1e-12 and ln.weight is not None
              and f.untyped_storage().nbytes() - f.storage_offset() * 4 >= -(-f.shape[0] // 128) * 128 * 163 * 4)
        if not ok:
            res = self.samvit_mlp(f)
            if nchw:
                return res.t().contiguous()
            if out is not None:
                out.view(-1, res.shape[-1]).copy_(res)
                return out.view(-1, res.shape[-1])
            return res
        lib = _lib.load()
        n, device = f.shape[0], f.device
        if getattr(self, "_sam_ws", None) is None or self._sam_ws.device != device:
            self._sam_ws = torch.empty(lib.sanerf_samvit_mlp_workspace_bytes(), dtype=torch.uint8, device=device)
        ws = [l.weight.detach().contiguous() for l in net]
        bs = [l.bias.detach().contiguous() for l in net]
        if out is None:
            out = torch.empty(n, 256, device=device)
        elif out.numel() != n * 256 or out.dtype != torch.float32 or out.device != device or not out.is_contiguous():
            raise RuntimeError(f"NeRFRenderer: out['samvit'] must be a contiguous fp32 tensor with {n} x 256 elements on {device}")
        else:
            out = out.view(n, 256)
        if nchw:
            out = out.view(256, n)
        wp = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in ws])
        bp = (ctypes.c_void_p * 5)(*[t.data_ptr() for t in bs])
        with torch.cuda.device(device), _lib.timed("samvit_mlp_kernel"):
            _lib.check(lib.sanerf_samvit_mlp_layout(_lib.ptr(f), wp, bp, _lib.ptr(ln.weight.detach()), _lib.ptr(ln.bias.detach()), n,
                                                    _lib.ptr(self._sam_ws), _lib.ptr(out), 1 if nchw else 0, _lib.stream_ptr()),
                       "sanerf_samvit_mlp")
            _lib.count_launch(2)
        return out

    # ------------------------------------------------------------------------------------------
    # composed path (differentiable: training)
    # ------------------------------------------------------------------------------------------
    def _run_composed(self, rays_o, rays_d, bg_color=None, perturb=False, cam_near_far=None, update_proposal=True,
                      return_feats=0, return_mask=0, H=None, W=None, feature_layout=None, feature_size=None, **kwargs):
        rays_o = rays_o.contiguous()
        rays_d = rays_d.contiguous()
        if not rays_o.is_cuda:
            raise RuntimeError("NeRFRenderer.run: rays must be CUDA tensors (there is no CPU path)")
        N, device = rays_o.shape[0], rays_o.device
        opt = self.opt

        near, far = near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer, self.min_near)
        if cam_near_far is not None:
            near = torch.maximum(near, cam_near_far[:, [0]])
            far = torch.minimum(far, cam_near_far[:, [1]])
        if bg_color is None:
            bg_color = 1
        s_near, s_far = _spacing(near), _spacing(far)

        collect = self.training
        all_bins, all_weights = [], []
        results = {}
        bins = weights = None
        n_stage = len(opt.num_steps)
        for stage, T in enumerate(opt.num_steps):
            if stage == 0:
                bins = torch.linspace(0, 1, T + 1, device=device).unsqueeze(0).expand(N, -1)
                if perturb:
                    bins = (bins + (torch.rand_like(bins) - 0.5) / T).clamp(0, 1)
            else:
                bins = sample_pdf(bins, weights, T + 1, perturb).detach()
            real_bins = _spacing_inv(s_near * (1 - bins) + s_far * bins)          # [N, T+1] in [near, far]
            rays_t = (real_bins[..., 1:] + real_bins[..., :-1]) / 2               # [N, T]
            xyzs = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * rays_t.unsqueeze(2)
            if opt.contract:
                xyzs = contract(xyzs)

            if stage != n_stage - 1:
                with torch.set_grad_enabled(update_proposal and torch.is_grad_enabled()):
                    sigmas = self.density(xyzs, proposal=stage)["sigma"]
            else:
                dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
                dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
                out = self(xyzs, dirs)
                sigmas, colors, geo_feat = out["sigma"], out["color"], out["geo_feat"]
                if opt.with_sam:
                    features = self.s_grid(xyzs, bound=self.bound)
                if return_mask > 0 and opt.mask_mlp_type in ("default", "lightweight_mask"):
                    masks = self.m_grid(xyzs, bound=self.bound)

            deltas = real_bins[..., 1:] - real_bins[..., :-1]
            ds = deltas * sigmas
            if opt.background == "last_sample":                                   # opaque far plane
                ds = torch.cat([ds[..., :-1], torch.full_like(ds[..., -1:], torch.inf)], dim=-1)
            alphas = 1 - torch.exp(-ds)
            acc = torch.cumsum(ds[..., :-1], dim=-1)
            trans = torch.exp(-torch.cat([torch.zeros_like(acc[..., :1]), acc], dim=-1))
            weights = (alphas * trans).nan_to_num_(0)
            if collect:
                all_bins.append(bins)
                all_weights.append(weights)

        weights_sum = torch.sum(weights, dim=-1)
        depth = torch.sum(weights * rays_t, dim=-1)
        f_image = torch.sum(weights.unsqueeze(-1) * colors, dim=-2)
        image = torch.sigmoid(self.view_mlp(f_image))

        if self.training and (not opt.with_mask and not opt.with_sam):
            results["num_points"] = xyzs.shape[0] * xyzs.shape[1]
            results["weights"] = weights
            if opt.lambda_proposal > 0 and update_proposal:
                results["proposal_loss"] = proposal_loss(all_bins, all_weights)
            if opt.lambda_distort > 0:
                results["distort_loss"] = distort_loss(bins, weights)

        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results["weights_sum"] = weights_sum
        results["depth"] = depth
        results["image"] = image

        if opt.with_sam:
            f_sam = torch.sum(weights.unsqueeze(-1) * features, dim=-2)
            if opt.sam_use_view_direction:
                f = torch.cat([f_sam, f_image, image, depth.unsqueeze(-1)], dim=-1)
            else:
                geo_sum = torch.sum(weights.unsqueeze(-1) * geo_feat, dim=-2)
                f = torch.cat([f_sam, geo_sum, image, depth.unsqueeze(-1)], dim=-1)
            samvit = self._samvit_head(f)
            if return_feats > 0:
                if feature_layout == "nchw":
                    nchw = samvit.view(1, H, W, -1).permute(0, 3, 1, 2).contiguous()
                    if feature_size is not None and tuple(feature_size) != (H, W):
                        nchw = torch.nn.functional.interpolate(nchw, tuple(feature_size), mode="bilinear")
                    results["samvit_nchw"] = nchw
                else:
                    results["samvit"] = samvit.view(H, W, -1)

        if return_mask > 0:
            if opt.mask_mlp_type == "default":
                point_masks = self.mask_mlp(torch.cat([masks, geo_feat.detach()], dim=-1))
            elif opt.mask_mlp_type == "lightweight_mask":
                point_masks = self.mask_mlp(torch.cat([masks, colors.detach()], dim=-1))
            results["instance_mask_logits"] = torch.sum(weights.detach().unsqueeze(-1) * point_masks, dim=-2)
        return results
