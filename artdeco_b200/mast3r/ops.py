"""Thin wrappers over the C ABI for the MASt3R path: tensor-core GEMM (bf16 / bf16x3) and its memory-bound
companions.  A "split" activation is a pair of bf16 tensors (hi, lo) with x ~= hi + lo; every producer kernel writes
the pair directly so GEMM operands never make a separate conversion pass."""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import f32, i32, i64, vp

_lib.register("adb_gemm_bf16", [i32, i32, i32, i32, vp, vp, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64,
                                vp, vp, i64, i64, f32, i32, i32, i64, i64, i64, vp])
_lib.register("adb_layernorm", [i64, i32, vp, vp, vp, f32, vp, vp, vp, vp])
_lib.register("adb_split_bf16", [i64, vp, vp, vp, vp])
_lib.register("adb_rope_heads", [i32, i32, i32, i64, i32, vp, vp, f32, i32, i32, vp, vp, vp])
_lib.register("adb_softmax_rows", [i64, i32, i64, i64, vp, vp, vp, vp])
_lib.register("adb_im2col_patch16", [i32, i32, i32, vp, vp, vp, vp])

BF16 = torch.bfloat16


class Split:
    """bf16 (hi, lo) pair; lo is None in single-pass bf16 mode."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape


def split(x: torch.Tensor, x3: bool = True) -> Split:
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=BF16, device=x.device)
    lo = torch.empty(x.shape, dtype=BF16, device=x.device) if x3 else None
    _lib.call("adb_split_bf16", x.numel(), _lib.ptr(x, torch.float32), _lib.ptr(hi), _lib.ptr(lo), _lib.stream())
    return Split(hi, lo)


def gemm(a: Split, w: Split, M: int, N: int, K: int, *, batch: int = 1, lda=None, sA=0, ldb=None, sB=0,
         out: torch.Tensor | None = None, ldd=None, sD=0, out_split: Split | None = None, ldo=None, sO=0,
         bias=None, residual=None, ldr=None, sR=0, alpha: float = 1.0, act: int = 0,
         zdiv: int = 0, sD2=0, sO2=0, sR2=0) -> None:
    """D = act(alpha * A @ W^T + bias) + residual, A: [batch, M, K], W: [batch, N, K] (both K-contiguous bf16)."""
    lda = K if lda is None else lda
    ldb = K if ldb is None else ldb
    ldd = N if ldd is None else ldd
    ldo = N if ldo is None else ldo
    ldr = N if ldr is None else ldr
    if batch > 1 and sB == 0:
        raise ValueError("broadcast weights over a batch: fold the batch into M")
    x3 = a.lo is not None and w.lo is not None
    _lib.call("adb_gemm_bf16", batch, M, N, K, _lib.ptr(a.hi), _lib.ptr(a.lo) if x3 else None, lda, sA,
              _lib.ptr(w.hi), _lib.ptr(w.lo) if x3 else None, ldb, sB,
              _lib.ptr(out) if out is not None else None, ldd, sD,
              _lib.ptr(out_split.hi) if out_split is not None else None,
              _lib.ptr(out_split.lo) if (out_split is not None and out_split.lo is not None) else None, ldo, sO,
              _lib.ptr(bias) if bias is not None else None,
              _lib.ptr(residual) if residual is not None else None, ldr, sR, float(alpha), int(act),
              int(zdiv), sD2, sO2, sR2, _lib.stream())


def linear(a: Split, w: Split, bias, rows: int, *, act: int = 0, residual=None, want_fp32=True, want_split=False,
           x3=True):
    """nn.Linear on a [rows, K] split activation; returns (fp32 [rows, N] or None, Split or None)."""
    N, K = w.hi.shape
    dev = a.hi.device
    out = torch.empty(rows, N, dtype=torch.float32, device=dev) if want_fp32 else None
    sp = None
    if want_split:
        sp = Split(torch.empty(rows, N, dtype=BF16, device=dev), torch.empty(rows, N, dtype=BF16, device=dev) if x3 else None)
    gemm(a, w, rows, N, K, out=out, out_split=sp, bias=bias, residual=residual, act=act)
    return out, sp


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-6, want_fp32=False, want_split=True, x3=True):
    C = x.shape[-1]
    rows = x.numel() // C
    x = x.contiguous()
    y = torch.empty_like(x) if want_fp32 else None
    sp = None
    if want_split:
        sp = Split(torch.empty(x.shape, dtype=BF16, device=x.device),
                   torch.empty(x.shape, dtype=BF16, device=x.device) if x3 else None)
    _lib.call("adb_layernorm", rows, C, _lib.ptr(x, torch.float32), _lib.ptr(gamma), _lib.ptr(beta), float(eps),
              _lib.ptr(y) if y is not None else None, _lib.ptr(sp.hi) if sp else None,
              _lib.ptr(sp.lo) if (sp and sp.lo is not None) else None, _lib.stream())
    return y, sp


def rope_heads(x: torch.Tensor, B, N, h, ld, col0, pos, mode: int, base: float = 100.0, x3=True, Npad=None) -> Split:
    dev = x.device
    Npad = N if Npad is None else Npad
    shape = (B, h, 64, Npad) if mode == 2 else (B, h, N, 64)
    alloc = torch.zeros if (mode == 2 and Npad != N) else torch.empty
    hi = alloc(shape, dtype=BF16, device=dev)
    lo = alloc(shape, dtype=BF16, device=dev) if x3 else None
    _lib.call("adb_rope_heads", B, N, h, ld, col0, _lib.ptr(x, torch.float32),
              _lib.ptr(pos, torch.int64) if pos is not None else None, float(base), mode, Npad, _lib.ptr(hi), _lib.ptr(lo),
              _lib.stream())
    return Split(hi, lo)


def softmax_rows(s: torch.Tensor, rows: int, L: int, ld_in: int, x3=True) -> Split:
    dev = s.device
    hi = torch.empty(rows, L, dtype=BF16, device=dev)
    lo = torch.empty(rows, L, dtype=BF16, device=dev) if x3 else None
    _lib.call("adb_softmax_rows", rows, L, ld_in, L, _lib.ptr(s, torch.float32), _lib.ptr(hi), _lib.ptr(lo), _lib.stream())
    return Split(hi, lo)


def im2col_patch16(img: torch.Tensor, x3=True) -> Split:
    B, C, H, W = img.shape
    assert C == 3
    n = (H // 16) * (W // 16)
    hi = torch.empty(B * n, 768, dtype=BF16, device=img.device)
    lo = torch.empty(B * n, 768, dtype=BF16, device=img.device) if x3 else None
    _lib.call("adb_im2col_patch16", B, H, W, _lib.ptr(img.contiguous(), torch.float32), _lib.ptr(hi), _lib.ptr(lo), _lib.stream())
    return Split(hi, lo)
