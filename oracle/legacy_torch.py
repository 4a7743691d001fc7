"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Pure-PyTorch, differentiable restatement of the LEGACY rasterizer contract the reference's web viewer uses
(Reconstruct/webviewer/scene_models.py:559-623: ``GaussianRasterizationSettings`` / ``GaussianRasterizer`` from the
on-the-fly-nvs fork of diff_gaussian_rasterization).  The fork is NOT vendored and no commit is pinned
(README.md:79), so this restates the published Inria algorithm: +0.3 px^2 low-pass, radius ceil(3 sqrt(lambda_max)) with a
0.1 floor under the root, near plane 0.2, tile rectangle (int)((p-r)/16)..(int)((p+r+15)/16) on pixel-index coordinates,
alpha = min(0.99, o G), skip alpha < 1/255, stop when T(1-alpha) < 1e-4, colour = sum c alpha T + T_final bg,
inverse depth = sum alpha T / z.  ``mainGaussID`` is taken to be the Gaussian with the largest weight alpha T.
PARITY UNPINNED: the reference holds no source, test or vector for this renderer.
Per-pixel Python loops: W*H <= ~64*64 and N <= ~200.
"""
from __future__ import annotations

import torch

from . import raster_torch as rt

MAX_ALPHA = 0.99
EPS2D = 0.3
NEAR = 0.2


def project(means, quats, scales, opacities, viewmat, K, W, H):
    """Returns radius[N] int32 (0 = culled / no tile), emit[N] bool, means2d, depths, conics."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    p = means @ R.T + t
    x, y, z = p.unbind(-1)
    M = rt.quat_to_rotmat(quats) * scales[:, None, :]
    Sc = R @ (M @ M.transpose(1, 2)) @ R.T
    tanx, tany = 0.5 * W / fx, 0.5 * H / fy
    lxp, lxn = (W - cx) / fx + 0.3 * tanx, cx / fx + 0.3 * tanx
    lyp, lyn = (H - cy) / fy + 0.3 * tany, cy / fy + 0.3 * tany
    zs = torch.where(z.abs() > 1e-12, z, torch.full_like(z, 1e-12))
    rz = 1.0 / zs
    tx = zs * torch.minimum(lxp, torch.maximum(-lxn, x * rz))
    ty = zs * torch.minimum(lyp, torch.maximum(-lyn, y * rz))
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], -1).reshape(-1, 2, 3)
    S2 = J @ Sc @ J.transpose(1, 2)
    means2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    a, b, c = S2[:, 0, 0] + EPS2D, S2[:, 0, 1], S2[:, 1, 1] + EPS2D
    det = a * c - b * b
    valid = (z >= NEAR) & (det > 0)
    dets = torch.where(det > 0, det, torch.ones_like(det))
    conics = torch.stack([c / dets, -b / dets, a / dets], -1)
    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam1 = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        radius = torch.ceil(3.0 * torch.sqrt(lam1)).to(torch.int32) * valid.to(torch.int32)
    return radius, means2d, z, conics


def _rect(means2d, radius, i, tw, th):
    r = float(radius[i])
    px, py = float(means2d[i, 0]) - 0.5, float(means2d[i, 1]) - 0.5
    f = lambda v: float(torch.tensor(v, dtype=torch.float32))
    x0 = min(tw, max(0, int(f(f(px - r) / 16.0))))
    x1 = min(tw, max(0, int(f(f(f(px + r) + 15.0) / 16.0))))
    y0 = min(th, max(0, int(f(f(py - r) / 16.0))))
    y1 = min(th, max(0, int(f(f(f(py + r) + 15.0) / 16.0))))
    return x0, x1, y0, y1


def rasterize(means, quats, scales, opacities, sh, viewmat, K, W, H, bg, sh_degree=3, scale_modifier=1.0):
    """Returns color[3,H,W], invdepth[1,H,W], mainGaussID[1,H,W] int32, radii[N] int32."""
    scales = scales * scale_modifier
    radius, means2d, depths, conics = project(means, quats, scales, opacities, viewmat, K, W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    radii_out = radius.clone()
    rects = {}
    for i in range(means.shape[0]):
        if int(radius[i]) <= 0:
            continue
        x0, x1, y0, y1 = _rect(means2d.detach(), radius, i, tw, th)
        if (x1 - x0) * (y1 - y0) == 0:
            radii_out[i] = 0                    # Inria: no tile touched -> radius stays 0
        elif float(opacities[i]) >= rt.ALPHA_THRESHOLD:
            rects[i] = (x0, x1, y0, y1)         # below 1/255 no pixel can pass the alpha test: nothing to emit
    campos = torch.inverse(viewmat)[:3, 3]
    rgb = rt.sh_colors(means, campos, sh, sh_degree)
    feats = torch.cat([rgb, 1.0 / depths[:, None]], -1)
    radii2 = torch.stack([radii_out, radii_out], -1)
    _, vals, offsets = rt.isect_and_sort(radii2, means2d, depths, W, H, rect_fn=lambda i, tw_, th_: rects.get(i))
    main = torch.zeros(H, W, dtype=torch.int32)
    out, alpha, _ = rt.blend(means2d, conics, opacities, feats, vals, offsets, W, H, max_alpha=MAX_ALPHA,
                             strict_stop=True, main_ids=main)
    color = (out[..., :3] + (1 - alpha[..., None]) * bg.view(1, 1, 3)).permute(2, 0, 1)
    return color, out[..., 3][None], main[None], radii_out
