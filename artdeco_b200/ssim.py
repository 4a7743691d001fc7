"""Host-side mirror of the reference ``fused_ssim`` operator surface.

Reference: Reconstruct/submodules/fused-ssim/fused_ssim/__init__.py:8-42 (``FusedSSIMMap``,
``fused_ssim``) and ext.cpp:4-7 (``fusedssim``, ``fusedssim_backward``).  Same names, argument order
and return conventions; the compute goes through the C ABI (``adb_ssim_forward/backward``).
"""
from __future__ import annotations

import torch

from . import _lib

allowed_padding = ["same", "valid"]


def _dims(img: torch.Tensor):
    if img.dim() != 4:
        raise ValueError("fused_ssim expects [B, CH, H, W] tensors")
    return tuple(int(s) for s in img.shape)


def fusedssim(C1: float, C2: float, img1: torch.Tensor, img2: torch.Tensor, train: bool):
    """Reference ext.cpp ``fusedssim`` (ssim.cu:434-478): returns (map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
    derivative tensors are empty when ``train`` is False."""
    _lib.require_cuda(img1)
    img1 = img1.contiguous()
    img2 = img2.contiguous()
    B, CH, H, W = _dims(img1)
    if tuple(img2.shape) != (B, CH, H, W):
        raise ValueError("img1 and img2 must have the same shape")
    ssim_map = torch.empty_like(img1)
    if train:
        d1, d2, d3 = torch.empty_like(img1), torch.empty_like(img1), torch.empty_like(img1)
    else:
        d1 = d2 = d3 = torch.empty(0, dtype=img1.dtype, device=img1.device)
    with torch.cuda.device(img1.device):
        _lib.call("adb_ssim_forward", B, CH, H, W, C1, C2, _lib.ptr(img1, torch.float32),
                  _lib.ptr(img2, torch.float32), int(train), _lib.ptr(ssim_map),
                  _lib.ptr(d1) if train else None, _lib.ptr(d2) if train else None,
                  _lib.ptr(d3) if train else None, None, _lib.stream())
    return ssim_map, d1, d2, d3


def fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """Reference ext.cpp ``fusedssim_backward`` (ssim.cu:480-517): returns dL/dimg1."""
    _lib.require_cuda(img1)
    img1, img2, dL_dmap = img1.contiguous(), img2.contiguous(), dL_dmap.contiguous()
    B, CH, H, W = _dims(img1)
    out = torch.empty_like(img1)
    with torch.cuda.device(img1.device):
        _lib.call("adb_ssim_backward", B, CH, H, W, C1, C2, _lib.ptr(img1, torch.float32),
                  _lib.ptr(img2, torch.float32), _lib.ptr(dL_dmap, torch.float32), None, 1.0,
                  _lib.ptr(dm_dmu1.contiguous()), _lib.ptr(dm_dsigma1_sq.contiguous()),
                  _lib.ptr(dm_dsigma12.contiguous()), _lib.ptr(out), _lib.stream())
    return out


class FusedSSIMMap(torch.autograd.Function):
    """Same contract as the reference class of this name (fused_ssim/__init__.py:8-32)."""

    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, d1, d2, d3 = fusedssim(C1, C2, img1, img2, train)
        if padding == "valid":
            ssim_map = ssim_map[:, :, 5:-5, 5:-5]
        ctx.save_for_backward(img1.detach(), img2, d1, d2, d3)
        ctx.C1, ctx.C2, ctx.padding = C1, C2, padding
        return ssim_map

    @staticmethod
    def backward(ctx, opt_grad):
        img1, img2, d1, d2, d3 = ctx.saved_tensors
        dL_dmap = opt_grad
        if ctx.padding == "valid":
            dL_dmap = torch.zeros_like(img1)
            dL_dmap[:, :, 5:-5, 5:-5] = opt_grad
        grad = fusedssim_backward(ctx.C1, ctx.C2, img1, img2, dL_dmap, d1, d2, d3)
        return None, None, grad, None, None, None


class _FusedSSIMMean(torch.autograd.Function):
    """``FusedSSIMMap(...).mean()`` for padding="same" in two launches and no map round trip: the
    forward reduces the map in-kernel, the backward uses a uniform upstream gradient."""

    @staticmethod
    def forward(ctx, C1, C2, img1, img2, train):
        _lib.require_cuda(img1)
        if tuple(img2.shape) != tuple(img1.shape):
            raise ValueError(f"fused_ssim: img1 {tuple(img1.shape)} and img2 {tuple(img2.shape)} must have the same shape")
        if not img2.is_cuda or img2.device != img1.device:
            raise ValueError(f"fused_ssim: img2 must be a CUDA tensor on {img1.device}, got {img2.device}")
        img2 = img2.contiguous()
        B, CH, H, W = _dims(img1)
        acc = torch.zeros(1, dtype=torch.float32, device=img1.device)
        if train:
            d1, d2, d3 = torch.empty_like(img1), torch.empty_like(img1), torch.empty_like(img1)
        else:
            d1 = d2 = d3 = None
        with torch.cuda.device(img1.device):
            _lib.call("adb_ssim_forward", B, CH, H, W, C1, C2, _lib.ptr(img1, torch.float32),
                      _lib.ptr(img2, torch.float32), int(train), None, _lib.ptr(d1), _lib.ptr(d2),
                      _lib.ptr(d3), _lib.ptr(acc), _lib.stream())
        if train:
            ctx.save_for_backward(img1.detach(), img2, d1, d2, d3)
        ctx.C1, ctx.C2, ctx.numel, ctx.train = C1, C2, img1.numel(), train
        return (acc / max(img1.numel(), 1)).reshape(())

    @staticmethod
    def backward(ctx, grad):
        if not ctx.train:
            raise RuntimeError("fused_ssim(train=False) is not differentiable")
        img1, img2, d1, d2, d3 = ctx.saved_tensors
        B, CH, H, W = _dims(img1)
        out = torch.empty_like(img1)
        g = grad.detach().to(torch.float32).reshape(1).contiguous()  # stays on the device: no host sync
        with torch.cuda.device(img1.device):
            _lib.call("adb_ssim_backward", B, CH, H, W, ctx.C1, ctx.C2, _lib.ptr(img1), _lib.ptr(img2),
                      None, _lib.ptr(g), 1.0 / ctx.numel, _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(d3), _lib.ptr(out), _lib.stream())
        return None, None, out, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    """Drop-in for ``fused_ssim.fused_ssim`` (fused_ssim/__init__.py:34-42): mean SSIM, grad to img1 only."""
    C1 = 0.01 ** 2
    C2 = 0.03 ** 2
    assert padding in allowed_padding
    img1 = img1.contiguous()
    if padding == "same":
        return _FusedSSIMMean.apply(C1, C2, img1, img2, train)
    return FusedSSIMMap.apply(C1, C2, img1, img2, padding, train).mean()
