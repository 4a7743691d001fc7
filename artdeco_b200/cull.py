"""LoD ``d_max`` cull of ``SceneModel.render`` (Reconstruct/scene/scene_models/h3dgsv3.py:626-645) as one fused
select + compaction launch, differentiable exactly where the reference's torch ops are (opacity, and xyz through
the fade ratio)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import f32, i32, i64, vp

_lib.register("adb_lod_select_workspace_bytes", [i64, C.POINTER(C.c_size_t)])
_lib.register("adb_lod_select", [i64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp])


_lib.register("adb_lod_weed_out", [i64, vp, vp, i32, vp, f32, vp, vp, vp])


def weed_out_mask(xyz: torch.Tensor, d_max: torch.Tensor, cam_centres: torch.Tensor, visible_threshold: float,
                  return_count: bool = False):
    """``weed_out_gaussians`` (h3dgsv3.py:942-953) without the per-key-frame Python loop: ``cam_centres`` [K,3] are the
    centres ``keyframe.get_Rt().T.inverse()[3,:3]`` of every key frame.  Returns the keep mask the reference hands to
    ``optimizer.add_and_prune(make_dummy_ext_tensor(), weed_mask)`` (and the per-Gaussian visible count)."""
    _lib.require_cuda(xyz)
    N = xyz.shape[0]
    dev = xyz.device
    cams = cam_centres.detach().float().reshape(-1, 3).contiguous().to(dev)
    if cams.shape[0] < 1:
        raise ValueError("weed_out_mask needs at least one key frame")
    keep = torch.empty(N, dtype=torch.bool, device=dev)
    cnt = torch.empty(N, dtype=torch.int32, device=dev) if return_count else None
    with torch.cuda.device(dev):
        _lib.call("adb_lod_weed_out", N, _lib.ptr(xyz.detach().float().contiguous()),
                  _lib.ptr(d_max.detach().float().reshape(-1).contiguous()), int(cams.shape[0]), _lib.ptr(cams),
                  float(visible_threshold), _lib.ptr(cnt), _lib.ptr(keep), _lib.stream())
    return (keep, cnt) if return_count else keep


def lod_select(xyz: torch.Tensor, d_max: torch.Tensor, cam_centre: torch.Tensor):
    """Returns (selection_mask[N0] bool, ids[N] int32 ascending, alpha_ratio[N0] fp32).  One host sync (count)."""
    _lib.require_cuda(xyz)
    N0 = xyz.shape[0]
    dev = xyz.device
    xyz_c = xyz.detach().float().contiguous()
    d = d_max.detach().float().reshape(-1).contiguous()
    cam = cam_centre.detach().float().reshape(3).contiguous()
    mask = torch.empty(N0, dtype=torch.bool, device=dev)
    ratio = torch.empty(N0, dtype=torch.float32, device=dev)
    ids = torch.empty(max(N0, 1), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    nb = C.c_size_t(0)
    _lib.call("adb_lod_select_workspace_bytes", N0, C.byref(nb))
    ws = torch.empty(max(nb.value, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("adb_lod_select", N0, _lib.ptr(xyz_c), _lib.ptr(d), _lib.ptr(cam), _lib.ptr(mask), _lib.ptr(ratio),
                  _lib.ptr(ids), _lib.ptr(count), _lib.ptr(ws), ws.numel(), _lib.stream())
    n = int(count.item())
    return mask, ids[:n], ratio


def lod_cull(xyz, d_max, opacity, cam_centre, *params):
    """The reference's cull block: returns (selection_mask, xyz_sel, opacity_sel * alpha_ratio, *[p[sel] for p in params]).
    Gradients flow to opacity and every gathered parameter, and to xyz through the fade ratio, as in the reference."""
    mask, ids, _ = lod_select(xyz, d_max, cam_centre)
    idx = ids.long()
    xyz_s = xyz.index_select(0, idx)
    d_s = d_max.reshape(-1, 1).index_select(0, idx)
    dist = (xyz_s - cam_centre.detach().reshape(1, 3)).norm(dim=1, keepdim=True)
    fade = (dist > d_s) & (dist < 2 * d_s)
    ratio = torch.where(fade, (2 * d_s - dist) / d_s, torch.ones_like(dist))
    op_s = opacity.reshape(xyz.shape[0], -1).index_select(0, idx) * ratio
    return (mask, xyz_s, op_s) + tuple(p.index_select(0, idx) for p in params)
