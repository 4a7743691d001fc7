"""Covariance-modulation MLP of ``SceneModel.render`` (Reconstruct/scene/scene_models/h3dgsv3.py:656-662)
as one fused forward and one fused backward kernel, with autograd to every tensor the reference's torch block reaches
(scaling, rotation, local_feat, global_feat, and the four mlp_cov parameters)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import i32, i64, vp

_lib.register("adb_cov_mlp_forward", [i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
_lib.register("adb_cov_mlp_backward", [i64, i32, i32] + [vp] * 20)


class _CovMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scaling, rotation, local_feat, global_feat, cls_id, W1, b1, W2, b2):
        _lib.require_cuda(scaling)
        N = scaling.shape[0]
        Fg, Fl = global_feat.shape[1], local_feat.shape[1]
        args = [t.detach().float().contiguous() for t in (scaling, rotation, local_feat, global_feat, W1, b1, W2, b2)]
        scaling_c, rotation_c, lf, gf, W1c, b1c, W2c, b2c = args
        cls = cls_id.reshape(-1).to(torch.int64).contiguous()
        s_out, r_out = torch.empty_like(scaling_c), torch.empty_like(rotation_c)
        with torch.cuda.device(scaling.device):
            _lib.call("adb_cov_mlp_forward", N, Fg, Fl, _lib.ptr(gf), _lib.ptr(lf), _lib.ptr(cls), _lib.ptr(W1c), _lib.ptr(b1c),
                      _lib.ptr(W2c), _lib.ptr(b2c), _lib.ptr(scaling_c), _lib.ptr(rotation_c), _lib.ptr(s_out), _lib.ptr(r_out),
                      _lib.stream())
        ctx.save_for_backward(scaling_c, rotation_c, lf, gf, cls, W1c, b1c, W2c, b2c)
        return s_out, r_out

    @staticmethod
    def backward(ctx, v_s, v_r):
        scaling, rotation, lf, gf, cls, W1, b1, W2, b2 = ctx.saved_tensors
        N, Fg, Fl = scaling.shape[0], gf.shape[1], lf.shape[1]
        v_s = torch.zeros_like(scaling) if v_s is None else v_s.float().contiguous()
        v_r = torch.zeros_like(rotation) if v_r is None else v_r.float().contiguous()
        g_s, g_r, g_lf = torch.empty_like(scaling), torch.empty_like(rotation), torch.empty_like(lf)
        g_gf, g_W1, g_b1, g_W2, g_b2 = (torch.zeros_like(t) for t in (gf, W1, b1, W2, b2))
        with torch.cuda.device(scaling.device):
            _lib.call("adb_cov_mlp_backward", N, Fg, Fl, _lib.ptr(gf), _lib.ptr(lf), _lib.ptr(cls), _lib.ptr(W1), _lib.ptr(b1),
                      _lib.ptr(W2), _lib.ptr(b2), _lib.ptr(scaling), _lib.ptr(rotation), _lib.ptr(v_s), _lib.ptr(v_r),
                      _lib.ptr(g_s), _lib.ptr(g_r), _lib.ptr(g_lf), _lib.ptr(g_gf), _lib.ptr(g_W1), _lib.ptr(g_b1),
                      _lib.ptr(g_W2), _lib.ptr(g_b2), _lib.stream())
        return g_s, g_r, g_lf, g_gf, None, g_W1, g_b1, g_W2, g_b2


def cov_mlp_modulate(scaling, rotation, local_feat, global_feat, cls_id, mlp_cov=None, *, W1=None, b1=None, W2=None, b2=None):
    """``scaling * sigmoid(o[:, :3]), normalize(rotation * o[:, 3:])`` with ``o = mlp_cov(cat(global_feat[cls_id], local_feat))``.
    ``mlp_cov`` may be the reference's ``nn.Sequential(Linear, ReLU, Linear)`` (its parameters are used and receive grads)."""
    if mlp_cov is not None:
        W1, b1, W2, b2 = mlp_cov[0].weight, mlp_cov[0].bias, mlp_cov[2].weight, mlp_cov[2].bias
    return _CovMlp.apply(scaling, rotation, local_feat, global_feat, cls_id, W1, b1, W2, b2)
