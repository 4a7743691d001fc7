"""Host-side mirror of ``gsplat.rendering.rasterization`` for the exact configuration the reference uses
(Reconstruct/scene/scene_models/h3dgsv3.py:664-680): 3DGS "classic" rasterisation, ``packed=False``,
``absgrad=False``, SH colours, ``render_mode`` "RGB" or "RGB+D".  Same keyword names, same return triple
``(render_colors[C,H,W,ch], render_alphas[C,H,W,1], meta)`` with ``meta['radii']`` of shape [C,N,2]
(h3dgsv3.py:689), and a complete autograd backward (means, quats, scales, opacities, colours, viewmats;
upstream grads for colours AND alphas, h3dgsv3.py:685-686).

All compute goes through the C ABI (adb_raster_*); PyTorch only owns memory, streams and the autograd graph.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import f32, i32, i64, vp

_lib.register("adb_raster_project_fwd", [i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, f32, f32, f32,
                                         vp, vp, vp, vp])
_lib.register("adb_raster_project_fwd_counts", [i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, f32, f32, f32,
                                                vp, vp, vp, vp, vp])
_lib.register("adb_raster_tile_scan", [i32, i32, i64, vp, vp, vp, vp, vp])
_lib.register("adb_raster_scan_workspace_bytes", [i32, C.POINTER(C.c_size_t)])
_lib.register("adb_raster_isect_scan", [i32, vp, vp, vp, C.c_size_t, vp])
_lib.register("adb_raster_isect_emit", [i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp])
_lib.register("adb_raster_sort_workspace_bytes", [i64, C.POINTER(C.c_size_t)])
_lib.register("adb_raster_sort", [i64, i32, i32, i32, vp, vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_int), vp])
_lib.register("adb_raster_tile_offsets", [i64, vp, i32, i32, vp, vp])
_lib.register("adb_raster_tile_count_scan", [i32, vp, vp, vp, i32, i32, i32, i64, vp, vp, vp, vp, vp])
_lib.register("adb_raster_tile_scatter_sort", [i32, vp, vp, vp, i32, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_blend_fwd", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_blend_bwd", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_project_bwd", [i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, f32, f32, f32,
                                         vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])

_lib.register("adb_raster_project_bwd_multi", [i32, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_sh_bwd_multi", [i32, i32, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp])
_lib.register("adb_raster_sh_dir_bwd_multi", [i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_sh_expand_multi", [i32, i32, vp, i32, vp, vp, vp, vp])

_lib.register("adb_raster_project_fwd_legacy", [i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, f32, f32,
                                                vp, vp, vp, vp])
_lib.register("adb_raster_isect_emit_legacy", [i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp])
_lib.register("adb_raster_blend_fwd_hits", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_blend_bwd_hits", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_blend_fwd_legacy", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_raster_blend_bwd_legacy", [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp])


TILE = 16
SPLAT_STRIDE = 12
COUNTER_COPIES = 4      # ADB_TILE_COUNTER_COPIES (csrc/raster_common.cuh)

_ws_cache: dict = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index, "ws")
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    return t


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def project(means, quats, scales, opacities, sh, sh_degree, viewmat, K, campos, W, H, eps2d, near, far, radius_clip,
            legacy=False, out=None, tile_counts=None):
    """Projection + SH + tile counts for one camera.  Returns radii[N,2] i32, splats[N,12], tiles_per_gauss[N].
    ``legacy``: Inria conventions (see include/artdeco_b200.h, adb_raster_project_fwd_legacy).
    ``out``: optional preallocated (radii, splats, tpg) — slices of per-camera stacks in the multi-view path.
    ``tile_counts``: optional zeroed counter buffer from ``new_tile_counts``: the per-tile counting of the bucketed
    intersection is fused into the projection kernel (pass the same buffer to ``intersect(counts=...)``)."""
    N = means.shape[0]
    dev = means.device
    if out is not None:
        radii, splats, tpg = out
    else:
        radii = torch.empty(N, 2, dtype=torch.int32, device=dev)
        splats = torch.empty(N, SPLAT_STRIDE, dtype=torch.float32, device=dev)
        tpg = torch.empty(N, dtype=torch.int32, device=dev)
    if legacy:
        _lib.call("adb_raster_project_fwd_legacy", N, _lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales),
                  _lib.ptr(opacities), _lib.ptr(sh), int(sh_degree), _lib.ptr(viewmat), _lib.ptr(K), _lib.ptr(campos),
                  W, H, eps2d, near, far, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(tpg), _lib.stream())
        return radii, splats, tpg
    if tile_counts is not None:
        _lib.call("adb_raster_project_fwd_counts", N, _lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(opacities),
                  _lib.ptr(sh), int(sh_degree), _lib.ptr(viewmat), _lib.ptr(K), _lib.ptr(campos), W, H, eps2d, near, far,
                  radius_clip, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(tpg), _lib.ptr(tile_counts), _lib.stream())
        return radii, splats, tpg
    _lib.call("adb_raster_project_fwd", N, _lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(opacities),
              _lib.ptr(sh), int(sh_degree), _lib.ptr(viewmat), _lib.ptr(K), _lib.ptr(campos), W, H, eps2d, near, far,
              radius_clip, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(tpg), _lib.stream())
    return radii, splats, tpg


def new_tile_counts(W, H, dev):
    """Zeroed counter buffer of the tile-bucketed intersection: 4 replicated counters per tile, then their segment starts."""
    T = ((W + TILE - 1) // TILE) * ((H + TILE - 1) // TILE)
    return torch.zeros(2 * COUNTER_COPIES * T, dtype=torch.int32, device=dev)


def intersect(radii, splats, tpg, W, H, cam_id=0, n_cams=1, sort=True, legacy=False, capacity=None, method=None, counts=None):
    """Tile keys/values (sorted), tile offsets [T+1] for one camera.

    ``method="bucket"`` (default): tile-bucketed pipeline — count, scan, scatter, one CTA-local bitonic sort per tile
    (csrc/raster_isect.cu) — bit-identical to the stable radix sort of (cam | tile | depth) keys it replaces.
    ``method="radix"`` (or env ADB_ISECT=radix): round-1 path (scan, emit, CUB onesweep radix sort, offsets), kept for A/B.

    ``capacity=None``: one host sync to size keys/vals exactly (gsplat does the same: ``cum_tiles[-1].item()``);
    returns ``(keys[I], vals[I], offsets, I:int)``.
    ``capacity=int``: NO host sync (CUDA-graph capturable): buffers hold ``capacity`` intersections, the true count and an
    overflow flag stay on the device; returns ``(keys[capacity], vals[capacity], offsets, info)`` with
    ``info = {"n_isect": int64[1] tensor, "overflow": int32[1] tensor}``; entries beyond ``offsets[T]`` are undefined.  On
    overflow the dropped intersections make the image wrong but nothing is written out of bounds.
    ``counts``: the buffer ``project(..., tile_counts=...)`` filled (counting fused into the projection): only the scan runs."""
    N = radii.shape[0]
    dev = radii.device
    T = ((W + TILE - 1) // TILE) * ((H + TILE - 1) // TILE)
    if method is None:
        import os
        method = os.environ.get("ADB_ISECT", "bucket")
    offsets = torch.empty(T + 1, dtype=torch.int32, device=dev)
    if N == 0:
        offsets.zero_()
        e64 = torch.empty(0, dtype=torch.int64, device=dev)
        e32 = torch.empty(0, dtype=torch.int32, device=dev)
        if capacity is None:
            return e64, e32, offsets, 0
        return e64, e32, offsets, {"n_isect": torch.zeros(1, dtype=torch.int64, device=dev),
                                   "overflow": torch.zeros(1, dtype=torch.int32, device=dev)}
    if method == "bucket" and sort:
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        cap_scan = int(capacity) if capacity is not None else 2147483646
        if counts is not None:
            _lib.call("adb_raster_tile_scan", W, H, cap_scan, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(total),
                      _lib.ptr(overflow), _lib.stream())
        else:
            counts = torch.zeros(2 * COUNTER_COPIES * T, dtype=torch.int32, device=dev)   # counters | segment starts
            _lib.call("adb_raster_tile_count_scan", N, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(tpg), W, H, int(legacy),
                      cap_scan, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(total), _lib.ptr(overflow), _lib.stream())
        if capacity is None:
            n_isect = int(total.item())      # the pipeline's single host sync
            if n_isect >= 2147483647:
                raise _lib.ArtdecoB200Error("more than 2^31 tile intersections")
            cap = n_isect
        else:
            cap = int(capacity)
        keys = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
        vals = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        if cap > 0:
            packed = torch.empty(cap, dtype=torch.int64, device=dev)
            _lib.call("adb_raster_tile_scatter_sort", N, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(tpg), W, H,
                      int(legacy), cam_id, n_cams, cap, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(packed),
                      _lib.ptr(keys), _lib.ptr(vals), _lib.stream())
        else:
            counts.zero_()
        if capacity is None:
            return keys[:n_isect], vals[:n_isect], offsets, n_isect
        return keys[:cap], vals[:cap], offsets, {"n_isect": total, "overflow": overflow}
    if capacity is not None:
        raise NotImplementedError("capacity mode needs method='bucket'")
    nb = C.c_size_t(0)
    _lib.call("adb_raster_scan_workspace_bytes", N, C.byref(nb))
    cum = torch.empty(N, dtype=torch.int64, device=dev)
    ws = _workspace(dev, nb.value)
    _lib.call("adb_raster_isect_scan", N, _lib.ptr(tpg), _lib.ptr(cum), _lib.ptr(ws), ws.numel(), _lib.stream())
    n_isect = int(cum[-1].item())  # the pipeline's single host sync
    keys_a = torch.empty(max(n_isect, 1), dtype=torch.int64, device=dev)
    vals_a = torch.empty(max(n_isect, 1), dtype=torch.int32, device=dev)
    if n_isect > 0:
        _lib.call("adb_raster_isect_emit_legacy" if legacy else "adb_raster_isect_emit", N, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(cum), W, H, cam_id, n_cams,
                  _lib.ptr(keys_a), _lib.ptr(vals_a), _lib.stream())
    keys, vals = keys_a, vals_a
    if sort and n_isect > 0:
        keys_b, vals_b = torch.empty_like(keys_a), torch.empty_like(vals_a)
        _lib.call("adb_raster_sort_workspace_bytes", n_isect, C.byref(nb))
        ws = _workspace(dev, nb.value)
        in_b = C.c_int(0)
        _lib.call("adb_raster_sort", n_isect, W, H, n_cams, _lib.ptr(keys_a), _lib.ptr(vals_a), _lib.ptr(keys_b),
                  _lib.ptr(vals_b), _lib.ptr(ws), ws.numel(), C.byref(in_b), _lib.stream())
        if in_b.value:
            keys, vals = keys_b, vals_b
    _lib.call("adb_raster_tile_offsets", n_isect, _lib.ptr(keys), W, H, _lib.ptr(offsets), _lib.stream())
    return keys[:n_isect], vals[:n_isect], offsets, n_isect


def new_hit_mask(vals: torch.Tensor):
    """One byte per sorted intersection for ``blend_forward(hits=)`` / ``blend_backward(hits=)``: the forward records which
    warps' pixel blocks each (tile, splat) entry can reach and the backward reuses those culling decisions instead of repeating
    the tests.  ``ADB_BLEND_HITS=0`` returns None (both kernels then cull on their own)."""
    if os.environ.get("ADB_BLEND_HITS", "1") == "0":
        return None
    return torch.empty(max(1, vals.numel()), dtype=torch.uint8, device=vals.device)


def blend_forward(W, H, N, splats, vals, offsets, legacy=False, out=None, hits=None):
    """``legacy`` returns a 4th tensor main_ids[H,W] and accumulates 1/z in colors[...,3].
    ``out``: optional preallocated (colors[H,W,4], alphas[H,W], last_ids[H,W]) — slices of per-camera stacks.
    ``hits``: optional uint8 [len(vals)] (``new_hit_mask``) that receives the culling decisions for ``blend_backward``."""
    dev = splats.device
    if out is not None:
        colors, alphas, last_ids = out
    else:
        colors = torch.empty(H, W, 4, dtype=torch.float32, device=dev)
        alphas = torch.empty(H, W, dtype=torch.float32, device=dev)
        last_ids = torch.empty(H, W, dtype=torch.int32, device=dev)
    if legacy:
        main_ids = torch.empty(H, W, dtype=torch.int32, device=dev)
        _lib.call("adb_raster_blend_fwd_legacy", W, H, N, _lib.ptr(splats), _lib.ptr(vals) if vals.numel() else None,
                  _lib.ptr(offsets), _lib.ptr(colors), _lib.ptr(alphas), _lib.ptr(last_ids), _lib.ptr(main_ids),
                  _lib.stream())
        return colors, alphas, last_ids, main_ids
    if hits is not None:
        if hits.dtype != torch.uint8 or hits.numel() < vals.numel() or hits.device != splats.device:
            raise ValueError("blend_forward: hits must be a uint8 tensor with one byte per intersection (new_hit_mask)")
        _lib.call("adb_raster_blend_fwd_hits", W, H, N, _lib.ptr(splats), _lib.ptr(vals) if vals.numel() else None,
                  _lib.ptr(offsets), _lib.ptr(colors), _lib.ptr(alphas), _lib.ptr(last_ids), _lib.ptr(hits), _lib.stream())
        return colors, alphas, last_ids
    _lib.call("adb_raster_blend_fwd", W, H, N, _lib.ptr(splats), _lib.ptr(vals) if vals.numel() else None,
              _lib.ptr(offsets), _lib.ptr(colors), _lib.ptr(alphas), _lib.ptr(last_ids), _lib.stream())
    return colors, alphas, last_ids


def blend_backward(W, H, N, splats, vals, offsets, alphas, last_ids, v_colors, v_alphas, legacy=False, out=None, hits=None):
    """``out``: optional ZEROED [N,12] accumulator (a slice of the per-camera stack in the multi-view path).
    ``hits``: the mask ``blend_forward(hits=)`` filled for the SAME splats / vals / offsets."""
    v_splats = out if out is not None else torch.zeros(N, SPLAT_STRIDE, dtype=torch.float32, device=splats.device)
    if hits is not None and not legacy:
        if hits.dtype != torch.uint8 or hits.numel() < vals.numel() or hits.device != splats.device:
            raise ValueError("blend_backward: hits must be the uint8 mask blend_forward filled for these intersections")
        _lib.call("adb_raster_blend_bwd_hits", W, H, N, _lib.ptr(splats), _lib.ptr(vals) if vals.numel() else None,
                  _lib.ptr(offsets), _lib.ptr(alphas), _lib.ptr(last_ids), _lib.ptr(v_colors), _lib.ptr(v_alphas),
                  _lib.ptr(v_splats), _lib.ptr(hits), _lib.stream())
        return v_splats
    _lib.call("adb_raster_blend_bwd_legacy" if legacy else "adb_raster_blend_bwd", W, H, N, _lib.ptr(splats), _lib.ptr(vals) if vals.numel() else None,
              _lib.ptr(offsets), _lib.ptr(alphas), _lib.ptr(last_ids), _lib.ptr(v_colors), _lib.ptr(v_alphas),
              _lib.ptr(v_splats), _lib.stream())
    return v_splats


class _RasterizeOneCamera(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, opacities, sh, colors_direct, viewmat, K, campos, W, H, sh_degree, eps2d,
                near, far, radius_clip, cam_id, n_cams):
        _lib.require_cuda(means)
        N = means.shape[0]
        with torch.cuda.device(means.device):
            cnt = new_tile_counts(W, H, means.device) if N > 0 else None
            radii, splats, tpg = project(means, quats, scales, opacities, sh, sh_degree, viewmat, K, campos, W, H,
                                         eps2d, near, far, radius_clip, tile_counts=cnt)
            if colors_direct is not None:
                splats[:, 8:11] = colors_direct
            keys, vals, offsets, n_isect = intersect(radii, splats, tpg, W, H, cam_id, n_cams, counts=cnt)
            hits = new_hit_mask(vals)
            colors, alphas, last_ids = blend_forward(W, H, N, splats, vals, offsets, hits=hits)
        ctx.hits = hits
        ctx.save_for_backward(means, quats, scales, opacities, sh, viewmat, K, campos, radii, splats, vals, offsets,
                              alphas, last_ids)
        ctx.cfg = (W, H, sh_degree, eps2d, near, far, radius_clip, colors_direct is not None)
        ctx.mark_non_differentiable(radii, splats, tpg, keys, vals, offsets)
        return colors, alphas, radii, splats, tpg, keys, vals, offsets

    @staticmethod
    def backward(ctx, v_colors, v_alphas, *_unused):
        (means, quats, scales, opacities, sh, viewmat, K, campos, radii, splats, vals, offsets, alphas,
         last_ids) = ctx.saved_tensors
        W, H, sh_degree, eps2d, near, far, radius_clip, direct = ctx.cfg
        N = means.shape[0]
        dev = means.device
        v_colors = torch.zeros(H, W, 4, device=dev) if v_colors is None else _f32c(v_colors)
        v_alphas = torch.zeros(H, W, device=dev) if v_alphas is None else _f32c(v_alphas)
        with torch.cuda.device(dev):
            v_splats = blend_backward(W, H, N, splats, vals, offsets, alphas, last_ids, v_colors, v_alphas, hits=ctx.hits)
            v_means = torch.empty_like(means)
            v_quats = torch.empty_like(quats)
            v_scales = torch.empty_like(scales)
            v_opac = torch.empty_like(opacities)
            v_sh = torch.empty_like(sh) if sh is not None else None
            v_view = torch.zeros(4, 4, dtype=torch.float32, device=dev)
            v_campos = torch.zeros(3, dtype=torch.float32, device=dev) if sh is not None else None
            _lib.call("adb_raster_project_bwd", N, _lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(sh),
                      int(sh_degree), _lib.ptr(viewmat), _lib.ptr(K), _lib.ptr(campos), W, H, eps2d, near, far,
                      radius_clip, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(v_splats), _lib.ptr(v_means),
                      _lib.ptr(v_quats), _lib.ptr(v_scales), _lib.ptr(v_opac), _lib.ptr(v_sh), _lib.ptr(v_view),
                      _lib.ptr(v_campos), _lib.stream())
        v_direct = None
        if direct:
            live = (radii > 0).any(-1, keepdim=True)
            v_direct = v_splats[:, 6:9] * live
        return (v_means, v_quats, v_scales, v_opac, v_sh, v_direct, v_view, None, v_campos,
                None, None, None, None, None, None, None, None, None)


class _RasterizeCameras(torch.autograd.Function):
    """C > 1 cameras of the same Gaussians (BASELINE config 4: an 8-view batch per optimiser step).  Forward: the
    single-view stages per camera into stacked buffers.  Backward: blend_bwd per camera, then ONE multi-view projection
    backward (parameters read once, 11 geometry gradients summed in registers) and ONE multi-view SH backward (v_sh
    written once) instead of C single-view passes plus C-1 [N,59] accumulations (csrc/raster_project_bwd_multi.cu)."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, sh, viewmats, Ks, camposs, W, H, sh_degree, eps2d, near, far,
                radius_clip, exchange=None, capacities=None):
        _lib.require_cuda(means)
        ctx.exchange = exchange
        N, Cn = means.shape[0], viewmats.shape[0]
        dev = means.device
        with torch.cuda.device(dev):
            radii = torch.empty(Cn, N, 2, dtype=torch.int32, device=dev)
            splats = torch.empty(Cn, N, SPLAT_STRIDE, dtype=torch.float32, device=dev)
            tpg = torch.empty(Cn, N, dtype=torch.int32, device=dev)
            colors = torch.empty(Cn, H, W, 4, dtype=torch.float32, device=dev)
            alphas = torch.empty(Cn, H, W, dtype=torch.float32, device=dev)
            last_ids = torch.empty(Cn, H, W, dtype=torch.int32, device=dev)
            keys_l, vals_l, offs_l, infos, hits_l = [], [], [], [], []
            for c in range(Cn):
                cnt = new_tile_counts(W, H, dev) if N > 0 else None
                project(means, quats, scales, opacities, sh, sh_degree, viewmats[c], Ks[c], camposs[c], W, H, eps2d, near,
                        far, radius_clip, out=(radii[c], splats[c], tpg[c]), tile_counts=cnt)
                cap = None if capacities is None else int(capacities[c] if hasattr(capacities, "__len__") else capacities)
                keys, vals, offsets, info = intersect(radii[c], splats[c], tpg[c], W, H, c, Cn, capacity=cap, counts=cnt)
                hits = new_hit_mask(vals)
                blend_forward(W, H, N, splats[c], vals, offsets, out=(colors[c], alphas[c], last_ids[c]), hits=hits)
                hits_l.append(hits)
                keys_l.append(keys)
                vals_l.append(vals)
                offs_l.append(offsets)
                infos.append(info)
        ctx.save_for_backward(means, quats, scales, opacities, sh, viewmats, Ks, camposs, radii, splats, alphas, last_ids,
                              *vals_l, *offs_l)
        ctx.cfg = (W, H, sh_degree, Cn)
        ctx.hits = hits_l
        isect_ids, flatten_ids = torch.cat(keys_l), torch.cat(vals_l)
        base, offs_out = 0, []
        for c in range(Cn):                        # gsplat offsets index the concatenated list (padded in capacity mode)
            offs_out.append(offs_l[c][:-1] + base)
            base += int(vals_l[c].numel())
        isect_offsets = torch.stack(offs_out)
        if capacities is None:
            overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            overflow = torch.stack([i["overflow"] for i in infos]).amax(0)
        ctx.mark_non_differentiable(radii, splats, tpg, isect_ids, flatten_ids, isect_offsets, overflow)
        return colors, alphas, radii, splats, tpg, isect_ids, flatten_ids, isect_offsets, overflow

    @staticmethod
    def backward(ctx, v_colors, v_alphas, *_unused):
        W, H, sh_degree, Cn = ctx.cfg
        saved = ctx.saved_tensors
        means, quats, scales, opacities, sh, viewmats, Ks, camposs, radii, splats, alphas, last_ids = saved[:12]
        vals_l, offs_l = saved[12:12 + Cn], saved[12 + Cn:12 + 2 * Cn]
        N = means.shape[0]
        dev = means.device
        v_colors = torch.zeros(Cn, H, W, 4, device=dev) if v_colors is None else _f32c(v_colors)
        v_alphas = torch.zeros(Cn, H, W, device=dev) if v_alphas is None else _f32c(v_alphas)
        with torch.cuda.device(dev):
            v_splats = torch.zeros(Cn, N, SPLAT_STRIDE, dtype=torch.float32, device=dev)
            for c in range(Cn):
                blend_backward(W, H, N, splats[c], vals_l[c], offs_l[c], alphas[c], last_ids[c], v_colors[c], v_alphas[c],
                               out=v_splats[c], hits=ctx.hits[c])
            ex = ctx.exchange
            grads = multi_view_backward(means, quats, scales, sh, sh_degree, viewmats, Ks, camposs, W, H, radii, splats,
                                        v_splats, out=(dict(ex.views) if ex is not None else None), exchange=ex)
        v_means, v_quats, v_scales, v_opac, v_sh, v_views, v_campos = grads
        if ex is not None:      # the bucket is reused by the next step: hand autograd its own copies
            v_means, v_quats, v_scales, v_opac = v_means.clone(), v_quats.clone(), v_scales.clone(), v_opac.clone()
        return (v_means, v_quats, v_scales, v_opac, v_sh, v_views, None, v_campos, None, None, None, None, None, None, None,
                None, None)


def mask_rgb_grad(splats, v_splats, out):
    """Colour gradient masked by the SH clamp (rgb = max(.,0): zero gradient where the record's colour is 0); invisible
    Gaussians have zero accumulators.  ``splats`` / ``v_splats`` [..., N, 12] -> ``out`` [..., N, 3]."""
    return torch.where(splats[..., 8:11] > 0, v_splats[..., 6:9], v_splats.new_zeros(()), out=out)


def multi_view_backward(means, quats, scales, sh, sh_degree, viewmats, Ks, camposs, W, H, radii, splats, v_splats,
                        out=None, exchange=None):
    """Projection + SH backward for C stacked local views (csrc/raster_project_bwd_multi.cu).

    ``out``: optional dict of preallocated gradient tensors (``v_means [N,3], v_quats [N,4], v_scales [N,3], v_opac [N],
    v_sh [N,16,3]``), e.g. views of a flat communication bucket.
    ``exchange`` (multi-GPU; ``peer.PeerExchange`` or ``parallel.MultiViewExchange``): the colour gradients of the local views
    are gathered while the geometry kernels run, the 11 geometry floats are summed over ranks while the SH gradient of EVERY
    rank's views is expanded; ``None`` = single process.  Returns (v_means, v_quats, v_scales, v_opac, v_sh, v_viewmats[C,4,4],
    v_campos[C,3] of the local views)."""
    N, Cn = means.shape[0], viewmats.shape[0]
    dev = means.device
    out = out or {}
    v_means = out.get("v_means") if out.get("v_means") is not None else torch.empty_like(means)
    v_quats = out.get("v_quats") if out.get("v_quats") is not None else torch.empty_like(quats)
    v_scales = out.get("v_scales") if out.get("v_scales") is not None else torch.empty_like(scales)
    v_opac = out.get("v_opac") if out.get("v_opac") is not None else torch.empty(N, dtype=torch.float32, device=dev)
    v_sh = out.get("v_sh") if out.get("v_sh") is not None else torch.empty_like(sh)
    v_views = torch.zeros(Cn, 4, 4, dtype=torch.float32, device=dev)
    v_campos = torch.zeros(Cn, 3, dtype=torch.float32, device=dev)
    Vc, Kc, Pc = _f32c(viewmats.detach()), _f32c(Ks.detach()), _f32c(camposs.detach())
    # colour gradient masked by the SH clamp (rgb = max(.,0): zero gradient where the record's colour is 0); invisible
    # Gaussians have zero accumulators.  Produced first so that its all-gather overlaps the geometry kernel.
    g_rgb = out.get("g_rgb")
    if exchange is not None and getattr(exchange, "gather_started", False):
        pass                                      # per-view gathers were issued as each view's blend backward finished
    else:
        if g_rgb is None:
            g_rgb = torch.empty(Cn, N, 3, dtype=torch.float32, device=dev)
        mask_rgb_grad(splats, v_splats, g_rgb)
        if exchange is not None:
            exchange.start_gather(g_rgb, Pc)
    # an exchange with separate input / result buffers (peer.PeerExchange): the kernel writes this rank's partial sums into
    # `views_in`, the reduced gradients appear in `views`
    gin = getattr(exchange, "views_in", None) if exchange is not None else None
    pm, pq, ps, po = (gin["v_means"], gin["v_quats"], gin["v_scales"], gin["v_opac"]) if gin is not None else \
        (v_means, v_quats, v_scales, v_opac)
    _lib.call("adb_raster_project_bwd_multi", N, Cn, _lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(Vc),
              _lib.ptr(Kc), W, H, _lib.ptr(radii), _lib.ptr(splats), _lib.ptr(v_splats), _lib.ptr(pm),
              _lib.ptr(pq), _lib.ptr(ps), _lib.ptr(po), None, _lib.ptr(v_views), _lib.stream())
    if gin is not None:
        gout = exchange.views
        v_means, v_quats, v_scales, v_opac = gout["v_means"], gout["v_quats"], gout["v_scales"], gout["v_opac"]
    if exchange is None:
        _lib.call("adb_raster_sh_bwd_multi", N, Cn, _lib.ptr(means), _lib.ptr(sh), int(sh_degree), _lib.ptr(Pc),
                  _lib.ptr(g_rgb), _lib.ptr(v_sh), _lib.ptr(v_means), 1, 0, 0, _lib.ptr(v_campos), _lib.stream())
        return v_means, v_quats, v_scales, v_opac, v_sh, v_views, v_campos
    # Multi-GPU: the SH backward in its split form.  The direction term (gradient of the colour w.r.t. the mean through the view
    # direction) is linear in the views, so each rank adds it for its LOCAL views to its partial geometry sums and the reduce
    # carries it; only the outer product basis(dir) (x) v_rgb needs every rank's colour gradients, and that expansion — the
    # part every rank repeats for all views — then runs without the SH coefficients and the basis gradients.
    _lib.call("adb_raster_sh_dir_bwd_multi", N, Cn, _lib.ptr(means), _lib.ptr(sh), int(sh_degree), _lib.ptr(Pc), _lib.ptr(splats),
              _lib.ptr(v_splats), _lib.ptr(pm), _lib.ptr(v_campos), _lib.stream())
    exchange.start_reduce()                       # geometry sums (v_means | v_quats | v_scales | v_opac), overlaps the expansion
    g_all, P_all = exchange.wait_gather()
    _lib.call("adb_raster_sh_expand_multi", N, int(g_all.shape[0]), _lib.ptr(means), int(sh_degree), _lib.ptr(P_all),
              _lib.ptr(g_all), _lib.ptr(v_sh), _lib.stream())
    exchange.wait_reduce()
    return v_means, v_quats, v_scales, v_opac, v_sh, v_views, v_campos


def rasterization(means: torch.Tensor, quats: torch.Tensor, scales: torch.Tensor, opacities: torch.Tensor,
                  colors: torch.Tensor, viewmats: torch.Tensor, Ks: torch.Tensor, width: int, height: int,
                  near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
                  sh_degree: Optional[int] = None, packed: bool = False, tile_size: int = 16,
                  backgrounds: Optional[torch.Tensor] = None, render_mode: str = "RGB",
                  rasterize_mode: str = "classic", absgrad: bool = False, grad_exchange=None, isect_capacity=None,
                  **unsupported):
    """See module docstring.  ``colors`` is SH [N,K,3] when ``sh_degree`` is given, else RGB [N,3].

    ``grad_exchange`` (not in gsplat; multi-GPU view-parallel training, BASELINE config 4): a
    ``parallel.MultiViewExchange``.  The backward then leaves on every rank the gradients summed over the views of ALL ranks
    (all-gather of the colour gradients + all-reduce of the geometry gradients inside the multi-view backward).

    ``isect_capacity`` (not in gsplat; int or one int per camera): size the intersection buffers by this bound instead of reading
    the count back — the call then has NO host sync (gsplat itself syncs on ``cum_tiles_per_gauss[-1].item()``).
    ``meta["isect_ids"] / ["flatten_ids"]`` are then padded to the capacity per camera and ``meta["isect_overflow"]`` (int32[1] on
    the device) is non-zero when a view needed more: read it together with the loss; the image of that view is then incomplete."""
    if unsupported:
        raise NotImplementedError(f"rasterization(): unsupported arguments {sorted(unsupported)}")
    if packed or absgrad or rasterize_mode != "classic" or tile_size != 16:
        raise NotImplementedError("only packed=False, absgrad=False, rasterize_mode='classic', tile_size=16 "
                                  "(the reference's configuration, h3dgsv3.py:664-680)")
    if render_mode not in ("RGB", "RGB+D"):
        raise NotImplementedError("render_mode must be 'RGB' or 'RGB+D'")
    _lib.require_cuda(means)
    N = means.shape[0]
    C_ = viewmats.shape[0]
    assert means.shape == (N, 3) and quats.shape == (N, 4) and scales.shape == (N, 3) and opacities.shape == (N,)
    assert viewmats.shape == (C_, 4, 4) and Ks.shape == (C_, 3, 3)
    means, quats, scales, opacities = _f32c(means), _f32c(quats), _f32c(scales), _f32c(opacities)
    sh = direct = None
    if sh_degree is not None:
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3
        if (sh_degree + 1) ** 2 > colors.shape[1]:
            raise ValueError("sh_degree needs more SH coefficients than colors provides")
        sh = _f32c(colors)
        if sh.shape[1] != 16:  # kernel reads a fixed 16x3 block
            sh = torch.cat([sh, sh.new_zeros(N, 16 - sh.shape[1], 3)], 1) if sh.shape[1] < 16 else sh[:, :16].contiguous()
    else:
        assert colors.shape == (N, 3)
        direct = _f32c(colors)
    th, tw = (height + TILE - 1) // TILE, (width + TILE - 1) // TILE
    if grad_exchange is not None and sh is None:
        raise NotImplementedError("grad_exchange needs SH colours (sh_degree given)")
    if isect_capacity is not None and sh is None:
        raise NotImplementedError("isect_capacity needs SH colours (sh_degree given)")
    if (C_ > 1 or grad_exchange is not None or isect_capacity is not None) and sh is not None:
        # multi-view batch: stacked buffers, one multi-view projection/SH backward (see _RasterizeCameras)
        Vs = _f32c(viewmats)
        camposs = torch.inverse(Vs)[:, :3, 3].contiguous()
        col, alp, radii, splats, tpg, isect_ids, flatten_ids, isect_offsets, overflow = _RasterizeCameras.apply(
            means, quats, scales, opacities, sh, Vs, _f32c(Ks).detach(), camposs, int(width), int(height), int(sh_degree),
            float(eps2d), float(near_plane), float(far_plane), float(radius_clip), grad_exchange, isect_capacity)
        if backgrounds is not None:
            col = torch.cat([col[..., :3] + (1.0 - alp[..., None]) * backgrounds.view(C_, 1, 1, 3), col[..., 3:]], -1)
        meta = {
            "radii": radii, "means2d": splats[..., 0:2], "depths": splats[..., 11], "conics": splats[..., 2:5],
            "opacities": opacities, "tiles_per_gauss": tpg, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
            "isect_offsets": isect_offsets.view(C_, th, tw), "isect_overflow": overflow,
            "width": width, "height": height, "tile_size": TILE, "tile_width": tw, "tile_height": th, "n_cameras": C_,
        }
        return (col if render_mode == "RGB+D" else col[..., :3]), alp[..., None], meta
    out_c, out_a, radii_l, metas = [], [], [], []
    base = 0
    for c in range(C_):
        V = _f32c(viewmats[c])
        K = _f32c(Ks[c]).detach()
        campos = torch.inverse(V)[:3, 3].contiguous() if sh is not None else None
        col, alp, radii, splats, tpg, keys, vals, offs = _RasterizeOneCamera.apply(
            means, quats, scales, opacities, sh, direct, V, K, campos, int(width), int(height),
            int(sh_degree) if sh_degree is not None else 0, float(eps2d), float(near_plane), float(far_plane),
            float(radius_clip), c, C_)
        if backgrounds is not None:
            col = torch.cat([col[..., :3] + (1.0 - alp[..., None]) * backgrounds[c].view(1, 1, 3), col[..., 3:]], -1)
        out_c.append(col if render_mode == "RGB+D" else col[..., :3])
        out_a.append(alp[..., None])
        radii_l.append(radii)
        metas.append((splats, tpg, keys, vals, offs + base))  # gsplat offsets index the concatenated list
        base += int(vals.numel())
    meta = {
        "radii": torch.stack(radii_l),
        "means2d": torch.stack([m[0][:, 0:2] for m in metas]),
        "depths": torch.stack([m[0][:, 11] for m in metas]),
        "conics": torch.stack([m[0][:, 2:5] for m in metas]),
        "opacities": opacities,
        "tiles_per_gauss": torch.stack([m[1] for m in metas]),
        "isect_ids": torch.cat([m[2] for m in metas]),
        "flatten_ids": torch.cat([m[3] for m in metas]),
        "isect_offsets": torch.stack([m[4][:-1].view(th, tw) for m in metas]),
        "width": width, "height": height, "tile_size": TILE, "tile_width": tw, "tile_height": th, "n_cameras": C_,
    }
    return torch.stack(out_c), torch.stack(out_a), meta
