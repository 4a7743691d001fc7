"""Host-side mirror of the reference's ``curope`` extension and its Python wrapper
(VSLAM/thirdparty/mast3r/dust3r/croco/models/curope/curope.cpp:49-68, kernels.cu:18-108, curope2d.py:12-40):
``rope_2d(tokens[B,N,H,D], positions[B,N,2], base, F0)`` rotates IN PLACE; ``cuRoPE2D(freq, F0)(tokens[B,H,N,D], positions)``.
The MASt3R forward of this package does not call it (RoPE is fused into the head-split kernel, adb_rope_heads); it exists
so that code written against the reference's operator (croco/models/pos_embed.py:106-109) keeps working.
fp32 CUDA tensors only — no CPU loop (the reference's rope_2d_cpu, curope.cpp:11-47, is not reproduced)."""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import f32, i32, i64, vp

_lib.register("adb_rope2d_inplace", [i32, i32, i32, i32, i64, i64, vp, vp, f32, f32, vp])


def rope_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float, F0: float) -> None:
    _lib.require_cuda(tokens)
    if tokens.dim() != 4 or positions.dim() != 3:
        raise ValueError("rope_2d: tokens must be [B,N,H,D] and positions [B,N,2]")
    B, N, H, D = tokens.shape
    if tokens.dtype != torch.float32:
        raise TypeError("rope_2d: fp32 tokens only")
    if positions.shape != (B, N, 2) or positions.dtype != torch.int64:
        raise ValueError("rope_2d: bad pos.shape / dtype (want int64 [B,N,2])")
    if D % 4 != 0:
        raise ValueError("rope_2d: token dim must be multiple of 4")
    if tokens.stride(3) != 1 or tokens.stride(2) != D:
        raise ValueError("rope_2d: tokens are not contiguous")
    positions = positions.contiguous()
    with torch.cuda.device(tokens.device):
        _lib.call("adb_rope2d_inplace", B, N, H, D, tokens.stride(0), tokens.stride(1), _lib.ptr(tokens),
                  _lib.ptr(positions), float(base), float(F0), _lib.stream())


class cuRoPE2D_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1):
        ctx.save_for_backward(positions)
        ctx.saved_base, ctx.saved_F0 = base, F0
        rope_2d(tokens, positions, base, F0)
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        rope_2d(grad_res, ctx.saved_tensors[0], ctx.saved_base, -ctx.saved_F0)
        ctx.mark_dirty(grad_res)
        return grad_res, None, None, None


class cuRoPE2D(torch.nn.Module):
    def __init__(self, freq=100.0, F0=1.0):
        super().__init__()
        self.base, self.F0 = freq, F0

    def forward(self, tokens, positions):
        cuRoPE2D_func.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens
