"""Host-side mirror of the reference's dense matching step:
``mast3r_slam_backends.iter_proj / refine_matches`` (VSLAM/backend/src/gn.cpp:84-112, matching_kernels.cu) and the functions
of ``VSLAM/utils_matching.py`` that call them — same names, argument meaning and return values.  ``match_iterative_proj``
runs four kernels (prep, LM projection, finalize, descriptor refinement) where the reference runs ~25 PyTorch ops around two.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32, i32, i64, vp

_lib.register("adb_match_prep", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_iter_proj", [i32, i32, i32, i32, vp, vp, vp, i32, f32, f32, vp, vp, vp])
_lib.register("adb_match_finalize", [i32, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp])
_lib.register("adb_refine_matches", [i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp, vp, vp])
_lib.register("adb_desc_pack_f16", [i32, i64, i32, vp, vp, vp])


def _chk(t, name, dtype):
    _lib.require_cuda(t)
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")      # gn.h:5 CHECK_CONTIGUOUS


def iter_proj(rays_img_with_grad, pts_3d_norm, p_init, max_iter, lambda_init, cost_thresh):
    """mast3r_slam_backends.iter_proj (gn.cpp:84-98): returns [p_new float [b,n,2], converged bool [b,n]]."""
    _chk(rays_img_with_grad, "rays_img_with_grad", torch.float32)
    _chk(pts_3d_norm, "pts_3d_norm", torch.float32)
    _chk(p_init, "p_init", torch.float32)
    b, h, w, c = rays_img_with_grad.shape
    if c != 9:
        raise ValueError("rays_img_with_grad must have 9 channels (ray, d/du, d/dv)")
    n = p_init.shape[1]
    p_new = torch.empty(b, n, 2, dtype=torch.float32, device=p_init.device)
    conv = torch.empty(b, n, dtype=torch.bool, device=p_init.device)
    with torch.cuda.device(p_init.device):
        _lib.call("adb_iter_proj", b, h, w, n, _lib.ptr(rays_img_with_grad), _lib.ptr(pts_3d_norm), _lib.ptr(p_init),
                  int(max_iter), float(lambda_init), float(cost_thresh), _lib.ptr(p_new), _lib.ptr(conv), _lib.stream())
    return [p_new, conv]


def refine_matches(D11, D21, p1, radius, dilation_max, return_linear: bool = False):
    """mast3r_slam_backends.refine_matches (gn.cpp:100-112): D11 [b,h,w,F], D21 [b,n,F] (fp16 as the reference passes them),
    p1 int64 [b,n,2] -> [p1_new]."""
    if D11.dtype != torch.float16 or D21.dtype != torch.float16:
        raise TypeError("refine_matches: descriptors must be fp16 (the reference calls it with .half(), utils_matching.py:178-184)")
    _chk(D11, "D11", torch.float16)
    _chk(D21, "D21", torch.float16)
    _chk(p1, "p1", torch.int64)
    b, h, w, F_ = D11.shape
    n = p1.shape[1]
    p1_new = torch.empty_like(p1)
    lin = torch.empty(b, n, dtype=torch.int64, device=p1.device) if return_linear else None
    with torch.cuda.device(p1.device):
        _lib.call("adb_refine_matches", b, h, w, F_, n, _lib.ptr(D11), _lib.ptr(D21), 0, _lib.ptr(p1), int(radius),
                  int(dilation_max), _lib.ptr(p1_new), _lib.ptr(lin), _lib.stream())
    return [p1_new, lin] if return_linear else [p1_new]


def _pack_f16(D, b, n_pix):
    """fp32 descriptors [b, n_pix, F] -> fp16 chunk-planar (adb_desc_pack_f16); same rounding as ``.half()``."""
    D = D.reshape(b, n_pix, -1)
    if D.dtype != torch.float32:
        D = D.float()
    D = D.contiguous()
    out = torch.empty(b * n_pix * D.shape[-1], dtype=torch.float16, device=D.device)
    _lib.call("adb_desc_pack_f16", b, n_pix, D.shape[-1], _lib.ptr(D), _lib.ptr(out), _lib.stream())
    return out, D.shape[-1]


def pixel_to_lin(p1, w):
    return p1[..., 0] + (w * p1[..., 1])


def lin_to_pixel(idx_1_to_2, w):
    return torch.stack((idx_1_to_2 % w, idx_1_to_2 // w), dim=-1)


def prep_for_iter_proj(X11, X21, idx_1_to_2_init):
    """utils_matching.py:120-145: returns rays_with_grad_img [b,h,w,9], pts3d_norm [b,hw,3], p_init [b,hw,2] (float)."""
    _chk(X11, "X11", torch.float32)
    X21 = X21.contiguous()
    _chk(X21, "X21", torch.float32)
    b, h, w, _ = X11.shape
    dev = X11.device
    rays = torch.empty(b, h, w, 9, dtype=torch.float32, device=dev)
    pts = torch.empty(b, h * w, 3, dtype=torch.float32, device=dev)
    p_init = torch.empty(b, h * w, 2, dtype=torch.float32, device=dev)
    if idx_1_to_2_init is not None:
        idx_1_to_2_init = idx_1_to_2_init.to(torch.int64).contiguous()
    with torch.cuda.device(dev):
        _lib.call("adb_match_prep", b, h, w, _lib.ptr(X11), _lib.ptr(X21), _lib.ptr(idx_1_to_2_init), _lib.ptr(rays),
                  _lib.ptr(pts), _lib.ptr(p_init), _lib.stream())
    return rays, pts, p_init


def _project_and_filter(cfg, X11, X21, idx_1_to_2_init):
    b, h, w = X21.shape[:3]
    X11 = X11.contiguous()
    X21 = X21.contiguous()
    rays, pts, p_init = prep_for_iter_proj(X11, X21, idx_1_to_2_init)
    p, conv = iter_proj(rays, pts, p_init, cfg["max_iter"], cfg["lambda_init"], cfg["convergence_thresh"])
    p1 = torch.empty(b, h * w, 2, dtype=torch.int64, device=X11.device)
    valid = torch.empty(b, h * w, dtype=torch.bool, device=X11.device)
    with torch.cuda.device(X11.device):
        _lib.call("adb_match_finalize", b, h, w, _lib.ptr(X11), _lib.ptr(X21), _lib.ptr(p), _lib.ptr(conv),
                  float(cfg["dist_thresh"]), _lib.ptr(p1), _lib.ptr(valid), _lib.stream())
    return p1, valid


def match_pi3(config, X11, X21, idx_1_to_2_init=None):
    """utils_matching.py:7-56 (projection + occlusion only)."""
    p1, valid = _project_and_filter(config["matching"], X11, X21, idx_1_to_2_init)
    return pixel_to_lin(p1, X21.shape[2]), valid


def match_iterative_proj(config, X11, X21, D11, D21, idx_1_to_2_init=None):
    """utils_matching.py:148-190: returns (idx_1_to_2 int64 [b,hw], valid_match2 bool [b,hw,1])."""
    cfg = config["matching"]
    b, h, w = X21.shape[:3]
    p1, valid = _project_and_filter(cfg, X11, X21, idx_1_to_2_init)
    if cfg["radius"] > 0:
        # fp32 -> fp16 conversion fused with the re-layout that makes the window gathers coalesce; identical arithmetic
        idx = torch.empty(b, h * w, dtype=torch.int64, device=p1.device)
        p1_new = torch.empty_like(p1)
        with torch.cuda.device(p1.device):
            a, F_ = _pack_f16(D11, b, h * w)
            q, _ = _pack_f16(D21, b, h * w)
            _lib.call("adb_refine_matches", b, h, w, F_, h * w, _lib.ptr(a), _lib.ptr(q), 1, _lib.ptr(p1),
                      int(cfg["radius"]), int(cfg["dilation_max"]), _lib.ptr(p1_new), _lib.ptr(idx), _lib.stream())
    else:
        idx = pixel_to_lin(p1, w)
    return idx, valid.unsqueeze(-1)


def match(config, X11, X21, D11, D21, idx_1_to_2_init=None):
    """utils_matching.py:99-101."""
    return match_iterative_proj(config, X11, X21, D11, D21, idx_1_to_2_init)
