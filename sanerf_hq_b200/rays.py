"""Pinhole ray generation for full frames -- the caller-side row f-2 of SURVEY.md 8f.

`get_rays` mirrors the full-image branch (N=-1) of the reference's nerf/utils.py::get_rays
(:183-304): pixel centres at +0.5, camera looks down -z with y up, `rays_d = dirs @ R^T` left
UNNORMALISED (the reference keeps it so for metric depth, utils.py:276-277), `rays_o = t`.
`orbit_pose` builds the look-at-origin cam2world matrices of the benchmark orbit.
Pure torch, any device.
"""
import math

import torch


def orbit_pose(k, n=24, radius=1.33, elev_deg=30.0, device="cpu"):
    az = 2 * math.pi * k / n
    el = math.radians(elev_deg)
    eye = torch.tensor([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az), radius * math.sin(el)],
                       dtype=torch.float64)
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, up, -fwd, eye
    return pose.float().to(device)


def lego_intrinsics(H, W, fov_x=0.6911112):
    """fx = fy = 0.5 W / tan(0.5 fov), principal point at the image centre (Blender-Lego camera)."""
    fl = 0.5 * W / math.tan(0.5 * fov_x)
    return fl, fl, W / 2, H / 2


def get_rays(pose, intrinsics, H, W, rows=None, device=None):
    """pose [4,4] cam2world, intrinsics (fx, fy, cx, cy) -> rays_o, rays_d [(r1-r0)*W, 3] for image rows
    [r0, r1) (default: the whole frame), row-major like the reference."""
    device = device or pose.device
    fx, fy, cx, cy = intrinsics
    r0, r1 = (0, H) if rows is None else rows
    j, i = torch.meshgrid(torch.arange(r0, r1, device=device, dtype=torch.float32),
                          torch.arange(0, W, device=device, dtype=torch.float32), indexing="ij")
    i = i.reshape(-1) + 0.5
    j = j.reshape(-1) + 0.5
    dirs = torch.stack(((i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)), dim=-1)
    rot = pose[:3, :3].to(device)
    rays_d = (dirs.unsqueeze(1) @ rot.t().unsqueeze(0)).squeeze(1)
    rays_o = pose[:3, 3].to(device).expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()
