"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Restatement of ``SparseGaussianAdam`` (Reconstruct/scene/optimizers.py:60-219) in plain PyTorch, device-agnostic:
  * ``adam``           the update of the un-vendored ``adamUpdate`` (no bias correction, invisible rows skipped);
  * ``step``           optimizers.py:76-161 for the per-Gaussian keys (update + per-primitive lr decay / clamp);
  * ``add_and_prune``  optimizers.py:163-219 literally (boolean indexing + cat + contiguous).
PARITY: add_and_prune is the reference's own PyTorch code path restated op for op (copies: exact); the Adam arithmetic is
unpinned by the reference (fork not vendored) and shared with oracle/raster_oracle.c::adbo_adam.
"""
from __future__ import annotations

import torch

NO_MOMENTS = ("id", "cls_id", "d_max")


def adam(param, grad, m, v, visible, lr, b1, b2, eps):
    vis = visible.view(-1, *([1] * (param.dim() - 1)))
    m_new = b1 * m + (1 - b1) * grad
    v_new = b2 * v + (1 - b2) * grad * grad
    p_new = param - lr * m_new / (v_new.sqrt() + eps)
    param.copy_(torch.where(vis, p_new, param))
    m.copy_(torch.where(vis, m_new, m))
    v.copy_(torch.where(vis, v_new, v))


def step(params, lr_dict, betas, eps, visibility):
    for key, pd in params.items():
        if key in NO_MOMENTS or key.startswith("mlp") or key == "global_feat":
            continue
        p = pd["val"]
        if p.grad is None:
            continue
        with torch.no_grad():
            adam(p, p.grad, pd["exp_avg"], pd["exp_avg_sq"], visibility, pd["lr"], betas[0], betas[1], eps)
            if key in lr_dict:
                pd["lr"][visibility] *= lr_dict[key]["lr_decay"]
                pd["lr"].clamp_min_(lr_dict[key]["lr_init"] * 0.1)


def add_and_prune(params, lr_dict, extension_tensors, valid_mask):
    for key, param in params.items():
        if key not in extension_tensors:
            continue
        ext = extension_tensors[key]
        empty = (ext.numel() == 0) or (ext.dim() == 0)
        if key == "global_feat":
            param["val"] = (param["val"].detach() if empty else torch.cat([param["val"].detach(), ext], 0)).contiguous()
            param["exp_avg"] = torch.cat([param["exp_avg"], torch.zeros_like(ext)], 0).contiguous()
            param["exp_avg_sq"] = torch.cat([param["exp_avg_sq"], torch.zeros_like(ext)], 0).contiguous()
            if key in lr_dict:
                param["lr"] = torch.cat([param["lr"], torch.ones_like(ext) * lr_dict[key]["lr_init"]], 0).contiguous()
            continue
        param["val"] = (param["val"].detach()[valid_mask] if empty
                        else torch.cat([param["val"].detach()[valid_mask], ext], 0)).contiguous()
        if key in NO_MOMENTS:
            continue
        param["val"].requires_grad = True
        param["exp_avg"] = torch.cat([param["exp_avg"][valid_mask], torch.zeros_like(ext)], 0).contiguous()
        param["exp_avg_sq"] = torch.cat([param["exp_avg_sq"][valid_mask], torch.zeros_like(ext)], 0).contiguous()
        if key in lr_dict:
            param["lr"] = torch.cat([param["lr"][valid_mask], torch.ones_like(ext) * lr_dict[key]["lr_init"]], 0).contiguous()
