"""CPU oracle of the render hot path: a functional torch restatement of
`NeRFRenderer.run` (nerf/renderer.py:221-385) + `NeRFNetwork.forward/density`
(nerf/network.py:146-188) over a plain dict of weights keyed by the reference's
state_dict names.  TEST INFRASTRUCTURE ONLY (see oracle/sanerf_oracle.c header).

The two CUDA-only encoders are evaluated by the C restatement (oracle/kernels.py);
everything else is the same sequence of torch CPU ops, in the same order and with
the same operand shapes as the reference, so that on the same torch build the
floating-point results coincide with the reference's Python run on CPU.  Pinned
against the reference itself by tests/golden/make_golden.py (config #1 fixtures).

Besides the reference's result dict, `run` returns the per-stage intermediates
the parity tests compare: bins, sigmas, weights and the sample_pdf index buffers.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import kernels as K


# ----------------------------------------------------------------------------------------
# model description (sizes hard-coded by nerf/network.py:90-144; overridable for config #1)
# ----------------------------------------------------------------------------------------

def grid_spec(num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
              desired_resolution=2048, input_dim=3):
    """Derived constants of one GridEncoder (gridencoder/grid.py:103-135)."""
    pls = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    return SimpleNamespace(num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                           log2_hashmap_size=log2_hashmap_size, per_level_scale=pls, input_dim=input_dim,
                           offsets=K.grid_offsets(input_dim, num_levels, base_resolution, pls, log2_hashmap_size))


def default_specs(bound=2, num_levels=None):
    """Grid specs of NeRFNetwork (network.py:93,102,120,136,141). `num_levels` shrinks every
    grid to that many levels (BASELINE config #1: L=4)."""
    L16 = num_levels or 16
    L5 = num_levels or 5
    return {
        "grid": grid_spec(L16, 2, 16, 19, 2048 * bound),
        "s_grid": grid_spec(L16, 8, 16, 19, 512),
        "m_grid": grid_spec(L16, 8, 16, 19, 512),
        "prop_encoders.0": grid_spec(L5, 2, 16, 17, 128),
        "prop_encoders.1": grid_spec(L5, 2, 16, 17, 256),
    }


def default_opt(**kw):
    """The 15 `opt` fields the hot path reads, with the values main.py
    produces for the shipped scripts (main.py:74-119, 217-221) and max_ray_batch per BASELINE."""
    o = dict(bound=128, contract=True, min_near=0.2, density_thresh=10, render_mesh=False,
             num_steps=[128, 64, 32], background="last_sample", with_sam=False, with_mask=False,
             mask_mlp_type="default", sam_use_view_direction=True, n_inst=2, max_ray_batch=4096,
             lambda_proposal=1, lambda_distort=0.02)
    o.update(kw)
    return SimpleNamespace(**o)


# ----------------------------------------------------------------------------------------
# network.py restated
# ----------------------------------------------------------------------------------------

def mlp_relu(x, weights):
    """MLP.forward (network.py:23-29): bias-free Linear stack, ReLU between layers."""
    for i, w in enumerate(weights):
        x = F.linear(x, w)
        if i != len(weights) - 1:
            x = F.relu(x)
    return x


def skip_mlp(x, weights, biases, skip_layers):
    """SkipConnMLP.forward (network.py:57-66): leaky_relu(0.01); at a skip layer the input is
    cat([hidden, x_in]) -- hidden first."""
    x_in = x
    n = len(weights)
    for l in range(n):
        if l in skip_layers:
            x = torch.cat([x, x_in], dim=-1)
        x = F.linear(x, weights[l], None if biases is None else biases[l])
        if l != n - 1:
            x = F.leaky_relu(x)
    return x


def _mlp_weights(params, prefix):
    ws, i = [], 0
    while f"{prefix}.net.{i}.weight" in params:
        ws.append(params[f"{prefix}.net.{i}.weight"])
        i += 1
    return ws


def encode_grid(params, specs, name, x, bound):
    sp = specs[name]
    return K.grid_encoder_apply(x, params[name + ".embeddings"], sp.offsets, sp.per_level_scale,
                                sp.base_resolution, bound=bound)


def density_proposal(params, specs, x, i, bound):
    """NeRFNetwork.density, proposal branch (network.py:173-178); trunc_exp fwd = exp (activation.py:10)."""
    h = encode_grid(params, specs, f"prop_encoders.{i}", x, bound)
    return torch.exp(mlp_relu(h, _mlp_weights(params, f"prop_mlp.{i}")).squeeze(-1))


def field_forward(params, specs, x, d, bound):
    """NeRFNetwork.forward (network.py:146-171)."""
    g = encode_grid(params, specs, "grid", x, bound)
    f = mlp_relu(g, _mlp_weights(params, "grid_mlp"))
    sigma = torch.exp(f[..., 0])
    feat = f[..., 1:]
    sh = K.sh_encoder_apply(d, degree=4)
    return sigma, feat, torch.cat([feat, sh], dim=-1), g


# ----------------------------------------------------------------------------------------
# renderer.py restated
# ----------------------------------------------------------------------------------------

def near_far_from_aabb(rays_o, rays_d, aabb, min_near):
    """renderer.py:122-139."""
    tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
    tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
    far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
    miss = far < near
    near = torch.where(miss, torch.full_like(near, 1e9), near)
    far = torch.where(miss, torch.full_like(far, 1e9), far)
    return near.clamp(min=min_near), far


def contract(x):
    """renderer.py:60-69: L-inf contraction to (-2, 2)."""
    shape = x.shape
    x = x.reshape(-1, shape[-1])
    mag, idx = x.abs().max(1, keepdim=True)
    scale = 1 / mag.repeat(1, shape[-1])
    scale.scatter_(1, idx, (2 - 1 / mag) / mag)
    return torch.where(mag < 1, x, x * scale).reshape(shape)


def sample_pdf(bins, weights, T, perturb=False, u_noise=None):
    """renderer.py:84-119.  Also returns the index buffers (inds, below, above) and cdf/u."""
    N, T0 = weights.shape
    weights = weights + 0.01
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1).clamp(max=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.5 / T, 1 - 0.5 / T, steps=T).expand(N, T)
    if perturb:
        u = u + ((torch.rand_like(u) if u_noise is None else u_noise) - 0.5) / T
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, 0, T0)
    above = torch.clamp(inds, 0, T0)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    t = torch.clamp(torch.nan_to_num((u - c0) / (c1 - c0)), 0, 1)
    return b0 + t * (b1 - b0), dict(inds=inds, below=below, above=above, cdf=cdf, u=u)


def _spacing(x):
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def _spacing_inv(x):
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


def weights_from_sigmas(real_bins, sigmas, last_sample):
    """renderer.py:308-325."""
    deltas = real_bins[..., 1:] - real_bins[..., :-1]
    ds = deltas * sigmas
    if last_sample:
        ds = torch.cat([ds[..., :-1], torch.full_like(ds[..., -1:], torch.inf)], dim=-1)
    alphas = 1 - torch.exp(-ds)
    tr = torch.cumsum(ds[..., :-1], dim=-1)
    tr = torch.exp(-torch.cat([torch.zeros_like(tr[..., :1]), tr], dim=-1))
    w = alphas * tr
    return torch.nan_to_num(w, 0)


@torch.no_grad()
def run(params, specs, opt, rays_o, rays_d, bg_color=None, cam_near_far=None, return_feats=0,
        return_mask=0, H=None, W=None, aabb=None):
    """Eval-mode `NeRFRenderer.run` with perturb=False.  CPU, fp32."""
    rays_o = rays_o.contiguous().float()
    rays_d = rays_d.contiguous().float()
    N = rays_o.shape[0]
    bound = 2 if opt.contract else opt.bound
    if aabb is None:
        aabb = params.get("aabb_infer", torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32))
    nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, opt.min_near)
    if cam_near_far is not None:
        nears = torch.maximum(nears, cam_near_far[:, [0]])
        fars = torch.minimum(fars, cam_near_far[:, [1]])
    if bg_color is None:
        bg_color = 1
    s_near, s_far = _spacing(nears), _spacing(fars)

    extras = dict(bins=[], real_bins=[], sigmas=[], weights=[], pdf=[])
    bins = weights = None
    n_stage = len(opt.num_steps)
    for it in range(n_stage):
        T = opt.num_steps[it]
        if it == 0:
            bins = torch.linspace(0, 1, T + 1).unsqueeze(0).expand(N, -1)
        else:
            bins, aux = sample_pdf(bins, weights, T + 1)
            extras["pdf"].append(aux)
        real_bins = _spacing_inv(s_near * (1 - bins) + s_far * bins)
        rays_t = (real_bins[..., 1:] + real_bins[..., :-1]) / 2
        xyzs = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * rays_t.unsqueeze(2)
        if opt.contract:
            xyzs = contract(xyzs)
        if it != n_stage - 1:
            sigmas = density_proposal(params, specs, xyzs, it, bound)
        else:
            dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
            dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
            sigmas, geo_feat, colors, _ = field_forward(params, specs, xyzs, dirs, bound)
            if opt.with_sam:
                features = encode_grid(params, specs, "s_grid", xyzs, bound)
            if return_mask > 0:
                masks = encode_grid(params, specs, "m_grid", xyzs, bound)
        weights = weights_from_sigmas(real_bins, sigmas, opt.background == "last_sample")
        extras["bins"].append(bins)
        extras["real_bins"].append(real_bins)
        extras["sigmas"].append(sigmas)
        extras["weights"].append(weights)

    weights_sum = torch.sum(weights, dim=-1)
    depth = torch.sum(weights * rays_t, dim=-1)
    f_image = torch.sum(weights.unsqueeze(-1) * colors, dim=-2)
    image = torch.sigmoid(mlp_relu(f_image, _mlp_weights(params, "view_mlp")))
    image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
    results = dict(weights_sum=weights_sum, depth=depth, image=image)
    extras["f_image"] = f_image

    if opt.with_sam:
        f_sam = torch.sum(weights.unsqueeze(-1) * features, dim=-2)
        assert opt.sam_use_view_direction, "only --sam_use_view_direction is functional (SURVEY.md section 0)"
        f = torch.cat([f_sam, f_image, image, depth.unsqueeze(-1)], dim=-1)
        ws = _mlp_weights(params, "samvit_mlp.0")
        bs = [params[f"samvit_mlp.0.net.{i}.bias"] for i in range(len(ws))]
        s = skip_mlp(f, ws, bs, skip_layers=[2])
        s = F.layer_norm(s, (s.shape[-1],), params["samvit_mlp.1.weight"], params["samvit_mlp.1.bias"], 1e-5)
        extras["f_samvit_in"] = f
        if return_feats > 0:
            results["samvit"] = s.view(H, W, -1)
    if return_mask > 0:
        m = torch.cat([masks, geo_feat], dim=-1)
        pm = skip_mlp(m, _mlp_weights(params, "mask_mlp.0"), None, skip_layers=[])
        results["instance_mask_logits"] = torch.sum(weights.unsqueeze(-1) * pm, dim=-2)
    return results, extras


@torch.no_grad()
def render(params, specs, opt, rays_o, rays_d, staged=True, **kw):
    """`NeRFRenderer.render` (renderer.py:185-219): chunk by opt.max_ray_batch and scatter."""
    if not staged:
        return run(params, specs, opt, rays_o, rays_d, **kw)[0]
    N = rays_o.shape[0]
    out = {}
    cnf = kw.pop("cam_near_far", None)
    for head in range(0, N, opt.max_ray_batch):
        tail = min(head + opt.max_ray_batch, N)
        c = None if cnf is None else (cnf if cnf.shape[0] == 1 else cnf[head:tail])
        r, _ = run(params, specs, opt, rays_o[head:tail], rays_d[head:tail], cam_near_far=c, **kw)
        for k, v in r.items():
            if k not in out:
                out[k] = torch.empty(N, *v.shape[1:])
            out[k][head:tail] = v
    return out


# ----------------------------------------------------------------------------------------
# synthetic weights / rays -- generator-seeded so the GPU box reproduces them
# ----------------------------------------------------------------------------------------

def _kaiming_linear(out_f, in_f, gen, bias=False):
    """nn.Linear default init (kaiming_uniform a=sqrt(5) -> U(-1/sqrt(in), 1/sqrt(in)))."""
    bnd = 1 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bnd
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bnd if bias else None
    return w, b


def make_params(opt, specs=None, seed=0, hidden=None, table_scale=1.0):
    """Random-init weights with the reference's state_dict key names and shapes.
    Hash tables ~ U(-1,1)*table_scale (the default U(-1e-4,1e-4) gives fog).
    `hidden` overrides the grid_mlp/view_mlp hidden width (config #1: 16)."""
    bound = 2 if opt.contract else opt.bound
    specs = specs or default_specs(bound)
    g = torch.Generator().manual_seed(seed)
    p = {}

    def table(name, k):
        sp = specs[name]
        tg = torch.Generator().manual_seed(1000 + k)
        p[name + ".embeddings"] = (torch.rand(int(sp.offsets[-1]), sp.level_dim, generator=tg) * 2 - 1) * table_scale
        p[name + ".offsets"] = sp.offsets

    def mlp(prefix, dims, bias=False):
        for i in range(len(dims) - 1):
            w, b = _kaiming_linear(dims[i + 1], dims[i], g, bias)
            p[f"{prefix}.net.{i}.weight"] = w
            if bias:
                p[f"{prefix}.net.{i}.bias"] = b

    gh = hidden or 64
    vh = hidden or 32
    table("grid", 0)
    mlp("grid_mlp", [specs["grid"].num_levels * 2, gh, gh, 16])
    mlp("view_mlp", [31, vh, vh, 3])
    for i in range(2):
        table(f"prop_encoders.{i}", 1 + i)
        mlp(f"prop_mlp.{i}", [specs[f"prop_encoders.{i}"].num_levels * 2, 16, 1])
    if opt.with_sam:
        table("s_grid", 3)
        din = specs["s_grid"].num_levels * 8 + 15 + 16 + 4
        for i, (fi, fo) in enumerate([(din, 256), (256, 256), (256 + din, 256), (256, 256), (256, 256)]):
            w, b = _kaiming_linear(fo, fi, g, True)
            p[f"samvit_mlp.0.net.{i}.weight"], p[f"samvit_mlp.0.net.{i}.bias"] = w, b
        p["samvit_mlp.1.weight"] = 1 + 0.1 * (torch.rand(256, generator=g) - 0.5)
        p["samvit_mlp.1.bias"] = 0.1 * (torch.rand(256, generator=g) - 0.5)
    if opt.with_mask:
        table("m_grid", 4)
        mlp("mask_mlp.0", [specs["m_grid"].num_levels * 8 + 15, 256, 256, opt.n_inst])
    p["aabb_train"] = torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32)
    p["aabb_infer"] = p["aabb_train"].clone()
    return p, specs


def orbit_pose(k, n=24, radius=1.33, elev_deg=30.0):
    """Look-at-origin cam2world on a seeded orbit (SURVEY.md 8d 'Rays'); OpenGL convention
    (camera looks down -z, y up), up = +z world."""
    az = 2 * math.pi * k / n
    el = math.radians(elev_deg)
    c = np.array([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az), radius * math.sin(el)])
    fwd = -c / np.linalg.norm(c)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, c
    return torch.from_numpy(m).float()


def get_rays(pose, H, W, fov_x=0.6911112):
    """Full-image branch of nerf/utils.py::get_rays (:183-304, N=-1): pixel centres +0.5,
    dirs = ((i-cx)/fx, -(j-cy)/fy, -1), rays_d = dirs @ R^T (unnormalised), rays_o = t."""
    fl = 0.5 * W / math.tan(0.5 * fov_x)
    cx, cy = W / 2, H / 2
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    i = i.reshape(-1) + 0.5
    j = j.reshape(-1) + 0.5
    dirs = torch.stack(((i - cx) / fl, -(j - cy) / fl, -torch.ones_like(i)), dim=-1)
    rays_d = (dirs.unsqueeze(1) @ pose[:3, :3].t().unsqueeze(0)).squeeze(1)
    rays_o = pose[:3, 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()
