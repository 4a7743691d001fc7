"""Seeded synthetic workloads.  Everything is generated on the CPU with a
``torch.Generator`` so the same tensors can be fed to the oracle and (after a host->device copy) to the
CUDA path; there is no dataset or checkpoint access."""
from __future__ import annotations

import math

import torch


def raster_scene(N: int, seed: int = 0, z_range=(2.0, 18.0), extent=(8.0, 4.5),
                 scale_range=(0.005, 0.05), dtype=torch.float32):
    """`raster_scene(N, seed)` of SURVEY.md §8d: world == camera-0 frame."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, 3, generator=g, dtype=dtype)
    means = torch.stack([(u[:, 0] * 2 - 1) * extent[0], (u[:, 1] * 2 - 1) * extent[1],
                         z_range[0] + u[:, 2] * (z_range[1] - z_range[0])], -1)
    ls = math.log(scale_range[0]) + torch.rand(N, 3, generator=g, dtype=dtype) * (
        math.log(scale_range[1]) - math.log(scale_range[0]))
    scales = torch.exp(ls)
    quats = torch.randn(N, 4, generator=g, dtype=dtype)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    opacities = torch.sigmoid(torch.randn(N, generator=g, dtype=dtype) * 1.5)
    sh = torch.cat([torch.randn(N, 1, 3, generator=g, dtype=dtype) * 0.5,
                    torch.randn(N, 15, 3, generator=g, dtype=dtype) * 0.05], 1)
    d_max = 4.0 + torch.rand(N, 1, generator=g, dtype=dtype) * 36.0
    return dict(means=means.contiguous(), quats=quats.contiguous(), scales=scales.contiguous(),
                opacities=opacities.contiguous(), sh=sh.contiguous(), d_max=d_max.contiguous())


def camera(width: int = 1920, height: int = 1080, view: float = 3.5, focal: float | None = None):
    """View v of SURVEY.md §8d: R = rot_y((v-3.5)*2deg), t = (0.2*(v-3.5), 0, 0); fx=fy=1000 at 1080p."""
    f = focal if focal is not None else 1000.0 * width / 1920.0
    K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
    ang = math.radians((view - 3.5) * 2.0)
    c, s = math.cos(ang), math.sin(ang)
    V = torch.eye(4, dtype=torch.float32)
    V[:3, :3] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    V[0, 3] = 0.2 * (view - 3.5)
    return V, K


def upstream_grads(width: int, height: int, seed: int = 1, C: int = 1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(C, height, width, 4, generator=g), torch.randn(C, height, width, 1, generator=g))


def ssim_pair(B=1, CH=3, H=1080, W=1920, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, CH, H, W, generator=g), torch.rand(B, CH, H, W, generator=g)
