"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

CPU/GPU-agnostic PyTorch restatement of the reference's dense matching step:
  * ``prep_for_iter_proj``  VSLAM/utils_matching.py:59-97,120-145 (this is the reference's own PyTorch code path, restated);
  * ``iter_proj``           VSLAM/backend/src/matching_kernels.cu:119-276, vectorised over points (fp32, the reference's
                            operation order; the reference build adds --use_fast_math, VSLAM/setup.py:62-67);
  * ``refine_matches``      matching_kernels.cu:26-83 with c10::Half semantics (product and running sum rounded to fp16,
                            sequential over the feature dimension; max_score starts at numeric_limits<Half>::min());
  * ``match_iterative_proj``utils_matching.py:148-190.
PINNED on the GPU box against the reference's OWN kernels: oracle/build_ref.py compiles matching_kernels.cu where it lies
into oracle/_ref/mast3r_matching_ref.so (tests/test_matching.py::test_matches_the_reference_cuda_build).  The reference
holds no tests or vectors for this path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def img_gradient(img):
    b, c, h, w = img.shape
    kx = (1.0 / 32.0) * torch.tensor([[-3.0, 0.0, 3.0], [-10.0, 0.0, 10.0], [-3.0, 0.0, 3.0]], dtype=img.dtype, device=img.device)
    ky = kx.t().contiguous()
    pad = F.pad(img, (1, 1, 1, 1), mode="reflect")
    return (F.conv2d(pad, kx.repeat(c, 1, 1, 1), groups=c), F.conv2d(pad, ky.repeat(c, 1, 1, 1), groups=c))


def prep_for_iter_proj(X11, X21, idx_init=None):
    b, h, w, _ = X11.shape
    rays = F.normalize(X11, dim=-1).permute(0, 3, 1, 2)
    gx, gy = img_gradient(rays)
    rays_grad = torch.cat((rays, gx, gy), 1).permute(0, 2, 3, 1).contiguous()
    pts = F.normalize(X21.reshape(b, -1, 3), dim=-1)
    if idx_init is None:
        idx_init = torch.arange(h * w, device=X11.device)[None].repeat(b, 1)
    p_init = torch.stack((idx_init % w, idx_init // w), -1).to(X11.dtype)
    return rays_grad, pts, p_init


def _bilinear(img, u, v):
    """img [b,h,w,c]; u,v [b,n] -> [b,n,c] with the reference's weights (matching_kernels.cu:153-170)."""
    b = img.shape[0]
    u11, v11 = torch.floor(u).long(), torch.floor(v).long()
    du, dv = u - u11.to(u.dtype), v - v11.to(v.dtype)
    bi = torch.arange(b, device=img.device)[:, None]
    w11, w12, w21, w22 = du * dv, (1 - du) * dv, du * (1 - dv), (1 - du) * (1 - dv)
    return (w11[..., None] * img[bi, v11 + 1, u11 + 1] + w12[..., None] * img[bi, v11 + 1, u11] +
            w21[..., None] * img[bi, v11, u11 + 1] + w22[..., None] * img[bi, v11, u11])


def iter_proj(rays_grad, pts, p_init, max_iter, lambda_init, cost_thresh):
    b, h, w, _ = rays_grad.shape
    u = p_init[..., 0].clamp(1, w - 2)
    v = p_init[..., 1].clamp(1, h - 2)
    lam = torch.full_like(u, lambda_init)
    conv = torch.zeros_like(u, dtype=torch.bool)
    for _ in range(max_iter):
        s = _bilinear(rays_grad, u, v)
        r, gx, gy = s[..., 0:3], s[..., 3:6], s[..., 6:9]
        r = r / r.norm(dim=-1, keepdim=True)
        err = r - pts
        cost = (err * err).sum(-1)
        A00 = (gx * gx).sum(-1) + lam
        A01 = (gx * gy).sum(-1)
        A11 = (gy * gy).sum(-1) + lam
        b0, b1 = -(err * gx).sum(-1), -(err * gy).sum(-1)
        det_inv = 1.0 / (A00 * A11 - A01 * A01)
        un = (u + det_inv * (A11 * b0 - A01 * b1)).clamp(1, w - 2)
        vn = (v + det_inv * (-A01 * b0 + A00 * b1)).clamp(1, h - 2)
        r2 = _bilinear(rays_grad[..., 0:3], un, vn)
        r2 = r2 / r2.norm(dim=-1, keepdim=True)
        new_cost = ((r2 - pts) ** 2).sum(-1)
        better = new_cost < cost
        u, v = torch.where(better, un, u), torch.where(better, vn, v)
        lam = torch.where(better, lam * 0.1, lam * 10.0)
        conv = torch.where(better, new_cost < cost_thresh, cost < cost_thresh)
    return torch.stack((u, v), -1), conv


def refine_matches(D11, D21, p1, radius, dilation_max):
    """D11 [b,h,w,F] fp16, D21 [b,n,F] fp16, p1 int64 [b,n,2].  fp16 arithmetic emulated step by step."""
    b, h, w, Fd = D11.shape
    n = p1.shape[1]
    dev = D11.device
    bi = torch.arange(b, device=dev)[:, None].expand(b, n)
    u0, v0 = p1[..., 0].clone(), p1[..., 1].clone()
    u_new, v_new = u0.clone(), v0.clone()
    max_score = torch.full((b, n), 6.103515625e-05, dtype=torch.float32, device=dev)
    q = D21.float()
    for d in range(dilation_max, 0, -1):
        rd = radius * d
        for i in range(0, 2 * rd + 1, d):
            for j in range(0, 2 * rd + 1, d):
                u, v = u0 - rd + i, v0 - rd + j
                ok = (v >= 0) & (v < h) & (u >= 0) & (u < w)
                c = D11[bi, v.clamp(0, h - 1), u.clamp(0, w - 1)].float()
                s = torch.zeros(b, n, dtype=torch.float32, device=dev)
                for k in range(Fd):
                    prod = (q[..., k] * c[..., k]).half().float()
                    s = (s + prod).half().float()
                upd = ok & (s > max_score)
                max_score = torch.where(upd, s, max_score)
                u_new, v_new = torch.where(upd, u, u_new), torch.where(upd, v, v_new)
        u0, v0 = u_new.clone(), v_new.clone()
    return torch.stack((u_new, v_new), -1)


def match_iterative_proj(cfg, X11, X21, D11, D21, idx_init=None):
    b, h, w = X21.shape[:3]
    rays_grad, pts, p_init = prep_for_iter_proj(X11, X21, idx_init)
    p, conv = iter_proj(rays_grad, pts, p_init, cfg["max_iter"], cfg["lambda_init"], cfg["convergence_thresh"])
    p1 = p.long()
    bi = torch.arange(b, device=X11.device)[:, None].repeat(1, h * w)
    d = torch.linalg.norm(X11[bi, p1[..., 1], p1[..., 0], :].reshape(b, h, w, 3) - X21, dim=-1)
    valid = conv & (d < cfg["dist_thresh"]).view(b, -1)
    if cfg["radius"] > 0:
        p1 = refine_matches(D11.half(), D21.reshape(b, h * w, -1).half(), p1, cfg["radius"], cfg["dilation_max"])
    return p1[..., 0] + w * p1[..., 1], valid.unsqueeze(-1), p, p1
