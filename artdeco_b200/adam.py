"""Host-side mirror of ``diff_gaussian_rasterization.adamUpdate`` / ``adamUpdateBasic`` (on-the-fly-nvs fork).

The fork is not vendored in the reference tree; the operator contract is pinned by its call sites
(Reconstruct/scene/optimizers.py:48-57, 90-99, 116-128, 144-156): positional arguments, in-place update of
``param``, ``exp_avg``, ``exp_avg_sq``, called under ``torch.no_grad()``, ``lr`` a 0-d / [N] / param-shaped CUDA
tensor (adamUpdate) or a Python float / tensor (adamUpdateBasic).  Adam WITHOUT bias correction.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import f32, i64, vp

_lib.register("adb_adam_update", [i64, i64, vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, vp])


def _launch(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, N, M):
    _lib.require_cuda(param)
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.ArtdecoB200Error(f"adamUpdate: {name} must be a contiguous float32 tensor")
    if param.numel() != N * M:
        raise ValueError(f"adamUpdate: param has {param.numel()} elements, expected N*M = {N * M}")
    lr_t, lr_s, lr_n = None, 0.0, 0
    if isinstance(lr, torch.Tensor):
        lr_t = lr.detach().to(device=param.device, dtype=torch.float32).contiguous()
        lr_n = lr_t.numel()
    else:
        lr_s = float(lr)
    vis = None
    if visible is not None:
        if visible.dtype not in (torch.bool, torch.uint8) or visible.numel() != N:
            raise ValueError("adamUpdate: visible must be a bool/uint8 tensor with N elements")
        vis = visible.contiguous()
    with torch.cuda.device(param.device):
        _lib.call("adb_adam_update", N, M, _lib.ptr(param), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                  _lib.ptr(vis), _lib.ptr(lr_t), lr_n, lr_s, float(b1), float(b2), float(eps), _lib.stream())


@torch.no_grad()
def adamUpdate(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, N, M):
    """In place; rows with ``visible[row] == False`` are skipped (moments untouched)."""
    _launch(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, int(N), int(M))


@torch.no_grad()
def adamUpdateBasic(param, grad, exp_avg, exp_avg_sq, lr, b1, b2, eps):
    """In place dense update of every element."""
    _launch(param, grad, exp_avg, exp_avg_sq, None, lr, b1, b2, eps, param.numel(), 1)
