"""Field network (reference nerf/network.py:9-231): hash grids + tiny MLPs.

Module / parameter names are part of the checkpoint boundary:
grid, grid_mlp.net.{0,1,2}, view_mlp.net.{0,1,2}, prop_encoders.{0,1}, prop_mlp.{0,1}.net.{0,1},
s_grid, samvit_mlp.0.net.{0..4}, samvit_mlp.1 (LayerNorm), m_grid, mask_mlp.0.net.{0,1,2}.
Construction order matches the reference so `torch.manual_seed(s); NeRFNetwork(opt)` consumes the
RNG identically.  Sizes are the reference's hard-coded ones; `num_levels` / `hidden_dim` exist only
to build BASELINE config #1 (every grid L=4, MLP width 16), which the reference cannot express.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .activation import trunc_exp
from .encoding import get_encoder
from .renderer import NeRFRenderer


class MLP(nn.Module):
    """Linear stack with in-place ReLU between layers, none after the last (network.py:9-29)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        widths = [dim_in] + [dim_hidden] * (num_layers - 1) + [dim_out]
        self.net = nn.ModuleList(nn.Linear(widths[i], widths[i + 1], bias=bias) for i in range(num_layers))

    def forward(self, x):
        last = self.num_layers - 1
        for i, layer in enumerate(self.net):
            x = layer(x)
            if i != last:
                x = F.relu(x, inplace=True)
        return x


class SkipConnMLP(nn.Module):
    """Leaky-ReLU MLP; at layer l in skip_layers the input is cat([hidden, x_in]) -- hidden first
    (network.py:31-66)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, skip_layers=[], bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden = dim_in, dim_out, dim_hidden
        self.num_layers, self.skip_layers = num_layers, skip_layers
        layers = []
        for l in range(num_layers):
            fan_in = dim_in if l == 0 else dim_hidden + (dim_in if l in skip_layers else 0)
            fan_out = dim_out if l == num_layers - 1 else dim_hidden
            layers.append(nn.Linear(fan_in, fan_out, bias=bias))
        self.net = nn.ModuleList(layers)

    def forward(self, x):
        x_in = x
        last = self.num_layers - 1
        for l, layer in enumerate(self.net):
            if l in self.skip_layers:
                x = torch.cat([x, x_in], dim=-1)
            x = layer(x)
            if l != last:
                x = F.leaky_relu(x, inplace=True)
        return x


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt, num_levels=None, hidden_dim=None):
        super().__init__(opt)
        L16, L5 = num_levels or 16, num_levels or 5
        self.geom_feat_dim = 15

        self.grid, self.grid_in_dim = get_encoder("hashgrid", input_dim=3, level_dim=2, num_levels=L16,
                                                  log2_hashmap_size=19, desired_resolution=2048 * self.bound)
        self.grid_mlp = MLP(self.grid_in_dim, 1 + self.geom_feat_dim, hidden_dim or 64, 3, bias=False)

        self.view_encoder, self.view_in_dim = get_encoder("sh", input_dim=3, degree=4)
        self.view_mlp = MLP(self.geom_feat_dim + self.view_in_dim, 3, hidden_dim or 32, 3, bias=False)

        if self.opt.with_sam:
            self.s_grid, self.s_dim = get_encoder("hashgrid", input_dim=3, num_levels=L16, level_dim=8,
                                                  base_resolution=16, log2_hashmap_size=19, desired_resolution=512)
            self.samvit_mlp_input_dim = self.s_dim + self.geom_feat_dim + 4
            if self.opt.sam_use_view_direction:
                self.samvit_mlp_input_dim += self.view_in_dim
            width = 256
            # NB the MLP is always built for the view-direction variant (network.py:114)
            self.samvit_mlp = nn.Sequential(
                SkipConnMLP(self.s_dim + self.geom_feat_dim + self.view_in_dim + 4, width, width, 5, skip_layers=[2], bias=True),
                nn.LayerNorm(width),
            )

        if self.opt.with_mask:
            if self.opt.mask_mlp_type == "default":
                self.m_grid, self.m_dim = get_encoder("hashgrid", input_dim=3, num_levels=L16, level_dim=8,
                                                      base_resolution=16, log2_hashmap_size=19, desired_resolution=512)
                self.mask_mlp = nn.Sequential(SkipConnMLP(self.m_dim + self.geom_feat_dim, self.opt.n_inst, 256, 3,
                                                          skip_layers=[], bias=False))
            elif self.opt.mask_mlp_type == "lightweight_mask":
                self.m_grid, self.m_dim = get_encoder("hashgrid", input_dim=3, num_levels=L16, level_dim=2,
                                                      base_resolution=16, log2_hashmap_size=10, desired_resolution=256)
                self.mask_mlp = MLP(self.geom_feat_dim + self.view_in_dim + 4, self.opt.n_inst, 64, 3, bias=False)

        # two proposal networks (network.py:131-144)
        self.prop_encoders = nn.ModuleList()
        self.prop_mlp = nn.ModuleList()
        for finest in (128, 256):
            enc, enc_dim = get_encoder("hashgrid", input_dim=3, level_dim=2, num_levels=L5, log2_hashmap_size=17,
                                       desired_resolution=finest)
            self.prop_encoders.append(enc)
            self.prop_mlp.append(MLP(enc_dim, 1, 16, 2, bias=False))

    def common_forward(self, x):
        grid_output = self.grid(x, bound=self.bound)
        f = self.grid_mlp(grid_output)
        return trunc_exp(f[..., 0]), f[..., 1:], grid_output

    def forward(self, x, d, **kwargs):
        # x [..., 3] in [-bound, bound]; d [..., 3] unit directions
        sigma, feat, grid_output = self.common_forward(x)
        return {"sigma": sigma, "geo_feat": feat, "color": torch.cat([feat, self.view_encoder(d)], dim=-1),
                "grid_output": grid_output}

    def density(self, x, proposal=-1):
        if 0 <= proposal < len(self.prop_encoders):
            h = self.prop_encoders[proposal](x, bound=self.bound)
            return {"sigma": trunc_exp(self.prop_mlp[proposal](h).squeeze(-1)), "geo_feat": None}
        sigma, feat, _ = self.common_forward(x)
        return {"sigma": sigma, "geo_feat": feat}

    def _regularised_grid(self):
        if self.opt.with_sam:
            return self.s_grid
        if self.opt.with_mask:
            return self.m_grid
        return self.grid

    def apply_total_variation(self, w):
        self._regularised_grid().grad_total_variation(w)

    def apply_weight_decay(self, w):
        self._regularised_grid().grad_weight_decay(w)

    def get_params(self, lr):
        groups = [self.grid, self.grid_mlp, self.view_mlp, self.prop_encoders, self.prop_mlp]
        if self.opt.with_sam:
            groups += [self.s_grid, self.samvit_mlp]
        if self.opt.with_mask:
            groups += [self.m_grid, self.mask_mlp]
        return [{"params": g.parameters(), "lr": lr} for g in groups]
