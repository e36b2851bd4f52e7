"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (its unmodified Python from
/root/reference) in the build container.  The fixtures pin the oracle (oracle/) and, through
it, the CUDA path.  Not run on the GPU box (no /root/reference there) -- the outputs are
committed.

How the reference is made to run on CPU without modification:
  * `mcubes`, `trimesh`, `torch_efficient_distloss` (dead imports on the eval path,
    nerf/renderer.py:6-7,14) are stubbed in sys.modules;
  * the CUDA-only pybind backends `_gridencoder` / `_shencoder` that
    gridencoder/grid.py:9-12 and shencoder/sphere_harmonics.py:9-12 import are provided as
    fake modules backed by the C restatement (oracle/sanerf_oracle.c).  Everything above
    them -- GridEncoder, SHEncoder, NeRFNetwork, NeRFRenderer.render/run, sample_pdf,
    contract, MLPs -- is the reference's own code.

Cases (weights from oracle.render_oracle.make_params, seeded; rays from the 8d orbit):
  cfg1_rgb   BASELINE config #1: 32x32 image, 16 rays/batch, every grid L=4, MLP hidden 16
  cfg1_sam   same + SAM feature head, one 8x8 low-res call (return_feats=1)
  cfg1_mask  same + object head (return_mask=1)
  full_rgb   default-size network (L=16/T=2^19, 2x64 MLP), 96 rays of the 800x800 frame
  full_sam   default sizes + SAM head, 5x8 rays
  full_mask  default sizes + object head, 64 rays
Option variants (small network; the option values travel inside the fixture as JSON):
  opt_white  --background white (no opaque last sample, the frame is mixed with bg_color)
  opt_box    contract=False, bound=1 (bounded scene: no contraction, aabb = [-1,1]^3)
  opt_cnf    per-ray cam_near_far through the staged render loop (renderer.py:197-205 slices it per chunk)
Pure-torch helpers (renderer.py:60-139):
  helpers    `contract` / `uncontract` / `near_far_from_aabb` / `sample_pdf` on adversarial inputs (ties, zeros, axis-parallel
             rays, empty and single-spike weight rows)
Training-only helpers (renderer.py:17-57):
  losses     the reference's `proposal_loss` and `distort_loss` on seeded three-stage bins / weights.  `distort_loss` calls the
             third-party `torch_efficient_distloss.eff_distloss` (requirements.txt:21, unpinned, not installed here): the fixture
             uses the PUBLISHED DEFINITION of that loss instead -- sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 delta_i, averaged
             over rays -- evaluated in float64 by brute force, plugged into the reference's own `distort_loss`.

    python tests/golden/make_golden.py [case ...]
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

# the reference's packages must win over this repo's drop-in packages of the same name
sys.path.insert(0, REF)
sys.path.append(REPO)

for name in ("mcubes", "trimesh", "torch_efficient_distloss"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["torch_efficient_distloss"].eff_distloss = None

from oracle import kernels as K  # noqa: E402
from oracle import render_oracle as O  # noqa: E402

_ge = types.ModuleType("_gridencoder")


def _gef(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype, align_corners, interp):
    K.grid_encode_forward(inputs, embeddings, offsets, B, D, C, L, max_level, S, H, dy_dx, gridtype,
                          align_corners, interp, outputs=outputs)


_ge.grid_encode_forward = _gef
sys.modules["_gridencoder"] = _ge
_sh = types.ModuleType("_shencoder")
_sh.sh_encode_forward = lambda inputs, outputs, B, D, C, dy_dx: K.sh_encode_forward(inputs, B, C, outputs=outputs)
sys.modules["_shencoder"] = _sh

import nerf.renderer as ref_renderer  # noqa: E402
import nerf.network as ref_network  # noqa: E402
from encoding import get_encoder  # noqa: E402

assert ref_renderer.__file__.startswith(REF) and ref_network.__file__.startswith(REF)


class SmallNetwork(ref_network.NeRFNetwork):
    """BASELINE config #1: the sizes are hard-coded in NeRFNetwork.__init__, so
    re-assign the modules after construction: every grid L=4, MLP hidden width 16."""

    def __init__(self, opt, L=4, hidden=16):
        super().__init__(opt)
        MLP, Skip = ref_network.MLP, ref_network.SkipConnMLP
        kw = dict(input_dim=3, num_levels=L)
        self.grid, d = get_encoder("hashgrid", level_dim=2, log2_hashmap_size=19, desired_resolution=2048 * self.bound, **kw)
        self.grid_mlp = MLP(d, 16, hidden, 3, bias=False)
        self.view_mlp = MLP(31, 3, hidden, 3, bias=False)
        if opt.with_sam:
            self.s_grid, sd = get_encoder("hashgrid", level_dim=8, log2_hashmap_size=19, desired_resolution=512, **kw)
            self.samvit_mlp = torch.nn.Sequential(Skip(sd + 15 + 16 + 4, 256, 256, 5, skip_layers=[2], bias=True),
                                                  torch.nn.LayerNorm(256))
        if opt.with_mask:
            self.m_grid, md = get_encoder("hashgrid", level_dim=8, log2_hashmap_size=19, desired_resolution=512, **kw)
            self.mask_mlp = torch.nn.Sequential(Skip(md + 15, opt.n_inst, 256, 3, skip_layers=[], bias=False))
        self.prop_encoders = torch.nn.ModuleList()
        self.prop_mlp = torch.nn.ModuleList()
        for des in (128, 256):
            e, pd = get_encoder("hashgrid", level_dim=2, log2_hashmap_size=17, desired_resolution=des, **kw)
            self.prop_encoders.append(e)
            self.prop_mlp.append(MLP(pd, 1, 16, 2, bias=False))


def capture_searchsorted():
    rec = []
    orig = torch.searchsorted

    def wrapped(*a, **k):
        out = orig(*a, **k)
        rec.append(out.clone())
        return out
    torch.searchsorted = wrapped
    return rec, lambda: setattr(torch, "searchsorted", orig)


def param_digest(params):
    return float(sum(float(v.double().abs().sum()) for k, v in sorted(params.items()) if v.is_floating_point()))


def make_case(name, small, with_sam=False, with_mask=False, H=32, W=32, rows=None, cols=None, batch=16, pose_k=3,
              return_feats=0, return_mask=0, staged=True, optkw=None, per_ray_near_far=False):
    torch.manual_seed(0)
    optkw = dict(optkw or {})
    opt = O.default_opt(with_sam=with_sam, with_mask=with_mask, max_ray_batch=batch, **optkw)
    specs = O.default_specs(2 if opt.contract else opt.bound, num_levels=4 if small else None)
    params, specs = O.make_params(opt, specs, seed=7, hidden=16 if small else None)
    model = (SmallNetwork(opt) if small else ref_network.NeRFNetwork(opt)).eval()
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys

    rays_o, rays_d = O.get_rays(O.orbit_pose(pose_k), H, W)
    if rows is not None:                      # sub-block of the frame
        idx = (torch.arange(rows[0], rows[1])[:, None] * W + torch.arange(cols[0], cols[1])[None, :]).reshape(-1)
        rays_o, rays_d = rays_o[idx].contiguous(), rays_d[idx].contiguous()
        h, w = rows[1] - rows[0], cols[1] - cols[0]
    else:
        h, w = H, W
    rec, restore = capture_searchsorted()
    with torch.no_grad():
        kw = dict(perturb=False, bg_color=1)
        if per_ray_near_far:                  # a different [near, far] window per ray, some of them empty (near > far)
            g = torch.Generator().manual_seed(5)
            near = 0.2 + 2.5 * torch.rand(rays_o.shape[0], generator=g)
            cnf = torch.stack([near, near + 4 * torch.rand(rays_o.shape[0], generator=g) - 0.3], dim=-1)
            kw["cam_near_far"] = cnf
        if return_feats:
            kw.update(return_feats=1, H=h, W=w)
        if return_mask:
            kw.update(return_mask=1)
        out = model.render(rays_o, rays_d, staged=staged, **kw)
    restore()
    n_chunks = len(rec) // 2
    inds0 = torch.cat(rec[0::2], 0)
    inds1 = torch.cat(rec[1::2], 0)
    assert inds0.shape == (rays_o.shape[0], 65) and inds1.shape == (rays_o.shape[0], 33), (inds0.shape, inds1.shape, n_chunks)

    # the oracle must agree with the reference it restates
    o_out = O.render(params, specs, opt, rays_o, rays_d, staged=staged,
                     **{k: v for k, v in kw.items() if k != "perturb"})
    for k in out:
        if torch.is_tensor(out[k]):
            a, b = out[k].reshape(-1), o_out[k].reshape(-1)
            err = ((a - b).abs() / b.abs().clamp(min=1e-3)).max().item()
            print(f"  {name}: oracle vs reference  {k:24s} max rel err {err:.3e}")
            assert err < 1e-5, (name, k, err)

    fix = dict(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), inds0=inds0.numpy().astype(np.int16),
               inds1=inds1.numpy().astype(np.int16), param_digest=np.float64(param_digest(params)),
               meta=np.array([int(small), int(with_sam), int(with_mask), h, w, batch, int(staged)], dtype=np.int32))
    if optkw:
        fix["optkw"] = np.array(json.dumps(optkw))
    if per_ray_near_far:
        fix["cam_near_far"] = kw["cam_near_far"].numpy()
    for k, v in out.items():
        if torch.is_tensor(v):
            fix["out_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(f"  wrote {name}.npz  ({rays_o.shape[0]} rays; keys {[k for k in fix if k.startswith('out_')]})")


def make_helpers(name="helpers"):
    """The reference's pure-torch helpers on adversarial inputs (renderer.py:60-139)."""
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(400, 3, generator=g) - 0.5) * 40
    x[:40] = (torch.rand(40, 3, generator=g) - 0.5) * 2                      # inside the unit cube: identity
    x[40] = torch.tensor([1.0, -1.0, 0.5])                                    # on the boundary, tie in the max coordinate
    x[41] = torch.tensor([3.0, 3.0, -3.0])                                    # three-way tie
    x[42] = torch.tensor([0.0, 0.0, 0.0])
    x[43] = torch.tensor([1e6, -2.0, 7.0])
    z = ref_renderer.contract(x)
    o = (torch.rand(300, 3, generator=g) - 0.5) * 6
    d = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1)
    d[:10, 0] = 0                                                             # axis-parallel: the +1e-15 guard
    d[10:20] = torch.tensor([0.0, 0.0, 1.0])
    o[20:40] *= 0.1                                                           # origins inside the box
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    near, far = ref_renderer.near_far_from_aabb(o, d, aabb, 0.2)
    fix = dict(contract_in=x.numpy(), contract_out=z.numpy(), uncontract_out=ref_renderer.uncontract(z).numpy(),
               rays_o=o.numpy(), rays_d=d.numpy(), aabb=aabb.numpy(), near=near.numpy(), far=far.numpy())
    for T0, T in ((128, 65), (64, 33)):
        N = 96
        w = torch.rand(N, T0, generator=g) ** 6
        w[:8] = 0                                                             # empty rays: uniform pdf
        w[8:16, : T0 // 2] = 0
        w[16, 5] = 1e30                                                       # one dominant sample
        w[17] = 1e-12
        edges = torch.sort(torch.rand(N, T0 + 1, generator=g), dim=-1).values
        edges[:, 0], edges[:, -1] = 0, 1
        edges[32:64] = torch.linspace(0, 1, T0 + 1)
        fix[f"pdf{T0}_bins"], fix[f"pdf{T0}_weights"] = edges.numpy(), w.numpy()
        fix[f"pdf{T0}_out"] = ref_renderer.sample_pdf(edges, w, T, perturb=False).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(f"  wrote {name}.npz  ({len(fix)} arrays)")


def make_losses(name="losses"):
    g = torch.Generator().manual_seed(3)
    N = 48
    all_bins, all_weights = [], []
    for T in (128, 64, 32):
        b = torch.sort(torch.rand(N, T + 1, generator=g), dim=-1).values
        b[:, 0], b[:, -1] = 0, 1
        w = torch.rand(N, T, generator=g) ** 4
        w = w / w.sum(-1, keepdim=True) * torch.rand(N, 1, generator=g)       # sums below one, like compositing weights
        all_bins.append(b)
        all_weights.append(w)
    all_weights[1][:3] = 0                                                     # empty rays

    def distloss_definition(w, m, interval):                                   # O(T^2), float64
        w, m, interval = w.double(), m.double(), interval.double()
        bi = (w[..., :, None] * w[..., None, :] * (m[..., :, None] - m[..., None, :]).abs()).sum(dim=(-1, -2))
        uni = (1 / 3) * (interval * w.pow(2)).sum(dim=-1)
        return (bi + uni).mean().float()

    ref_renderer.eff_distloss = distloss_definition
    pl = ref_renderer.proposal_loss(all_bins, all_weights)
    dl = ref_renderer.distort_loss(all_bins[-1], all_weights[-1])
    fix = {f"bins{i}": b.numpy() for i, b in enumerate(all_bins)}
    fix.update({f"weights{i}": w.numpy() for i, w in enumerate(all_weights)})
    fix.update(proposal_loss=np.float32(pl), distort_loss=np.float32(dl))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(f"  wrote {name}.npz  proposal_loss {float(pl):.6e}  distort_loss {float(dl):.6e}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    print("reference:", ref_renderer.__file__)
    cases = {
        "cfg1_rgb": dict(small=True),
        "cfg1_sam": dict(small=True, with_sam=True, H=8, W=8, batch=64, return_feats=1, staged=False),
        "cfg1_mask": dict(small=True, with_mask=True, H=16, W=16, batch=64, return_mask=1),
        "full_rgb": dict(small=False, H=800, W=800, rows=(396, 404), cols=(394, 406), batch=4096),
        "full_sam": dict(small=False, with_sam=True, H=800, W=800, rows=(300, 305), cols=(200, 208), batch=4096, return_feats=1,
                         staged=False),
        "full_mask": dict(small=False, with_mask=True, H=800, W=800, rows=(100, 108), cols=(600, 608), batch=4096, return_mask=1),
        "opt_white": dict(small=True, H=16, W=16, batch=64, optkw=dict(background="white")),
        "opt_box": dict(small=True, H=16, W=16, batch=64, optkw=dict(contract=False, bound=1)),
        "opt_cnf": dict(small=True, H=16, W=16, batch=64, per_ray_near_far=True),
    }
    for name in (sys.argv[1:] or list(cases) + ["losses", "helpers"]):      # `python make_golden.py opt_white opt_box` regenerates only those
        if name == "losses":
            make_losses()
        elif name == "helpers":
            make_helpers()
        else:
            make_case(name, **cases[name])
