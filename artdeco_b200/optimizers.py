"""Host-side mirror of ``Reconstruct/scene/optimizers.py`` (``BaseAdam``, ``SparseGaussianAdam``): the caller on the
optimiser side of the hot path.  Same constructor arguments, the same ``params`` dict-of-dicts
layout (``val`` / ``lr`` / ``exp_avg`` / ``exp_avg_sq``), the same ``step`` / ``add_and_prune`` semantics.

What runs differently:
  * ``step``: for per-primitive learning rates (keys in ``lr_dict``) the Adam update and the schedule
    ``lr[visible] *= lr_decay; lr.clamp_min_(0.1 lr_init)`` (optimizers.py:130-133,158-161: a masked index_put with a host sync
    plus a clamp) are ONE kernel, ``adb_adam_update_decay``;
  * ``add_and_prune``: every per-Gaussian tensor (params, both moments, learning rates, ids) is compacted and extended by ONE
    index plan + ONE gather launch (``adb_compact_plan`` / ``adb_compact_gather``) instead of boolean indexing + cat +
    contiguous per tensor (optimizers.py:163-219).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import f32, i32, i64, vp
from .adam import adamUpdate, adamUpdateBasic

_lib.register("adb_adam_update_decay", [i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, vp])
_lib.register("adb_compact_workspace_bytes", [i64, C.POINTER(C.c_size_t)])
_lib.register("adb_compact_plan", [i64, vp, vp, vp, vp, C.c_size_t, vp])
_lib.register("adb_compact_gather", [i64, i64, vp, i32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_uint), vp])

_NO_MOMENTS = ("id", "cls_id", "d_max")
MAX_TENSORS = 32


def _is_empty(t: torch.Tensor) -> bool:
    return t.numel() == 0 or t.dim() == 0


def _check_sparse_args(key, param, grad, m1, m2, visibility, N, M):
    """The raw-pointer launches below trust these shapes; a stale mask after densify must raise (the reference's boolean
    indexing raises IndexError in the same situation) instead of reading or writing out of bounds on the device."""
    if N * M != param.numel():
        raise ValueError(f"SparseGaussianAdam[{key}]: param has {param.numel()} elements, expected N*M = {N}*{M}")
    if visibility is None or visibility.dtype not in (torch.bool, torch.uint8) or visibility.numel() != N:
        got = None if visibility is None else (visibility.dtype, visibility.numel())
        raise IndexError(f"SparseGaussianAdam[{key}]: visibility must be a bool/uint8 tensor with N={N} elements, got {got}")
    for name, t in (("param", param), ("grad", grad), ("exp_avg", m1), ("exp_avg_sq", m2)):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != N * M or t.device != param.device:
            raise ValueError(f"SparseGaussianAdam[{key}]: {name} must be a contiguous float32 tensor of {N * M} elements "
                             f"on {param.device}")
    if visibility.device != param.device:
        raise ValueError(f"SparseGaussianAdam[{key}]: visibility is on {visibility.device}, params on {param.device}")


def compact_plan(valid_mask: torch.Tensor):
    """Returns (src_of int32 [N], n_keep).  One host sync (the count), like the first boolean index in the reference."""
    _lib.require_cuda(valid_mask)
    N = valid_mask.numel()
    dev = valid_mask.device
    mask = valid_mask.reshape(-1).contiguous()
    if mask.dtype not in (torch.bool, torch.uint8):
        raise TypeError("valid_mask must be bool/uint8")
    src_of = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    nb = C.c_size_t(0)
    _lib.call("adb_compact_workspace_bytes", N, C.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("adb_compact_plan", N, _lib.ptr(mask), _lib.ptr(src_of), _lib.ptr(cnt), _lib.ptr(ws), ws.numel(),
                  _lib.stream())
    return src_of, int(cnt.item())


def compact_gather(src_of, n_keep: int, n_ext: int, jobs, n_mask: int | None = None):
    """jobs: list of (src [N,...], ext [n_ext,...] | None, fill, tail_shape) -> list of outputs [n_keep+n_ext, *tail_shape].
    ``ext is None`` fills the n_ext tail rows with ``fill`` (no extension buffer is materialised)."""
    outs = []
    if n_mask is None:
        n_mask = int(src_of.numel())
    for j0 in range(0, len(jobs), MAX_TENSORS):
        chunk = jobs[j0:j0 + MAX_TENSORS]
        n = len(chunk)
        srcs, dsts, exts = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
        rws, fills = (C.c_int * n)(), (C.c_uint * n)()
        keep_alive = []
        for t, (src, ext, fill, tail) in enumerate(chunk):
            tail = tuple(tail)
            dtype = src.dtype
            if src.numel() and tuple(src.shape[1:]) != tail:
                raise ValueError(f"state rows {tuple(src.shape[1:])} do not match the extension rows {tail}")
            if src.numel() and src.shape[0] != n_mask:
                raise IndexError(f"compact_gather: state tensor has {src.shape[0]} rows but the mask has {n_mask} "
                                 "(stale valid_mask after a densification?)")
            if ext is not None and ext.numel() and (ext.shape[0] != n_ext or tuple(ext.shape[1:]) != tail):
                raise ValueError(f"compact_gather: extension shape {tuple(ext.shape)} != ({n_ext}, *{tail})")
            if dtype not in (torch.float32, torch.int64, torch.int32):
                raise TypeError(f"compact_gather: unsupported dtype {dtype}")
            src = src.contiguous()
            ext_c = ext.to(dtype).contiguous() if ext is not None else None
            out = torch.empty((n_keep + n_ext,) + tail, dtype=dtype, device=src.device)
            row_elems = 1
            for d in tail:
                row_elems *= int(d)
            words = row_elems * (2 if dtype == torch.int64 else 1)
            if dtype == torch.float32:
                fw = C.c_uint.from_buffer_copy(C.c_float(float(fill))).value
            else:
                if fill != 0:
                    raise ValueError("integer tensors can only be zero-filled")
                fw = 0
            srcs[t] = _lib.ptr(src).value if src.numel() else None
            dsts[t] = _lib.ptr(out).value if out.numel() else None
            exts[t] = _lib.ptr(ext_c).value if ext_c is not None and ext_c.numel() else None
            rws[t], fills[t] = max(words, 1), fw
            keep_alive += [src, ext_c]
            outs.append(out)
        if n_keep + n_ext > 0:
            with torch.cuda.device(outs[-1].device):
                _lib.call("adb_compact_gather", n_keep, n_ext, _lib.ptr(src_of), n, srcs, dsts, exts, rws, fills,
                          _lib.stream())
    return outs


class BaseAdam:
    """optimizers.py:17-58."""

    @torch.no_grad()
    def __init__(self, params, betas=(0.9, 0.999), eps=1e-15):
        self.params = params
        self.betas = betas
        self.eps = eps
        for param in self.params.values():
            if "exp_avg" not in param:
                param["exp_avg"] = torch.zeros_like(param["val"], memory_format=torch.preserve_format)
                param["exp_avg_sq"] = torch.zeros_like(param["val"], memory_format=torch.preserve_format)

    def zero_grad(self):
        for param in self.params.values():
            param["val"].grad = None

    @torch.no_grad()
    def step(self):
        for param_dict in self.params.values():
            param = param_dict["val"]
            if param.grad is None:
                continue
            adamUpdateBasic(param, param.grad, param_dict["exp_avg"], param_dict["exp_avg_sq"], param_dict["lr"],
                            self.betas[0], self.betas[1], self.eps)


class SparseGaussianAdam(BaseAdam):
    """optimizers.py:60-219."""

    def __init__(self, params, betas=(0.9, 0.999), eps=1e-15, lr_dict={}, device="cuda:0"):  # noqa: B006 (reference signature)
        super().__init__(params=params, betas=betas, eps=eps)
        self.device = device
        self.lr_dict = lr_dict
        for key, param in self.params.items():
            if key in _NO_MOMENTS or key.startswith("mlp"):
                continue
            if key not in self.lr_dict:
                param["lr"] = torch.tensor(param["lr"], dtype=torch.float, device=self.device)
            else:
                param["lr"] = torch.empty(0, dtype=torch.float, device=self.device)

    def _sparse_update(self, key, param_dict, visibility, N):
        param = param_dict["val"]
        if param.grad is None:
            return
        lr = param_dict["lr"]
        M = param.numel() // N
        if key in self.lr_dict and lr.numel() == param.numel() and lr.is_contiguous() and lr.dtype == torch.float32:
            cfg = self.lr_dict[key]
            grad = param.grad.contiguous()
            _check_sparse_args(key, param, grad, param_dict["exp_avg"], param_dict["exp_avg_sq"], visibility, N, M)
            vis = visibility.contiguous()
            with torch.cuda.device(param.device):
                _lib.call("adb_adam_update_decay", N, M, _lib.ptr(param), _lib.ptr(grad), _lib.ptr(param_dict["exp_avg"]),
                          _lib.ptr(param_dict["exp_avg_sq"]), _lib.ptr(vis), _lib.ptr(lr), float(self.betas[0]),
                          float(self.betas[1]), float(self.eps), float(cfg["lr_decay"]), float(cfg["lr_init"] * 0.1),
                          _lib.stream())
            return
        adamUpdate(param, param.grad, param_dict["exp_avg"], param_dict["exp_avg_sq"], visibility, lr, self.betas[0],
                   self.betas[1], self.eps, N, M)
        if key in self.lr_dict:
            param_dict["lr"][visibility] *= self.lr_dict[key]["lr_decay"]
            param_dict["lr"].clamp_min_(self.lr_dict[key]["lr_init"] * 0.1)

    @torch.no_grad()
    def step(self, visibility, N, global_visibility, N_global):
        for key, param_dict in self.params.items():
            if key in _NO_MOMENTS:
                continue
            if key.startswith("mlp"):
                param = param_dict["val"]
                if param.grad is None:
                    continue
                adamUpdateBasic(param, param.grad, param_dict["exp_avg"], param_dict["exp_avg_sq"], param_dict["lr"],
                                self.betas[0], self.betas[1], self.eps)
                if key in self.lr_dict:
                    param_dict["lr"] *= self.lr_dict[key]["lr_decay"]
                    param_dict["lr"] = max(param_dict["lr"], self.lr_dict[key]["lr_init"] * 0.1)
            elif key == "global_feat":
                self._sparse_update(key, param_dict, global_visibility, N_global)
            else:
                self._sparse_update(key, param_dict, visibility, N)

    @torch.no_grad()
    def add_and_prune(self, extension_tensors, valid_mask):
        # global_feat is append-only and tiny (one row per class): same ops as the reference (optimizers.py:169-194)
        if "global_feat" in extension_tensors and "global_feat" in self.params:
            param, ext = self.params["global_feat"], extension_tensors["global_feat"]
            if _is_empty(ext):
                param["val"] = param["val"].detach().contiguous()
            else:
                param["val"] = torch.cat([param["val"].detach(), ext], dim=0).contiguous()
            param["val"].requires_grad = True
            param["exp_avg"] = torch.cat([param["exp_avg"], torch.zeros_like(ext)], dim=0).contiguous()
            param["exp_avg_sq"] = torch.cat([param["exp_avg_sq"], torch.zeros_like(ext)], dim=0).contiguous()
            if "global_feat" in self.lr_dict:
                param["lr"] = torch.cat([param["lr"], torch.ones_like(ext) * self.lr_dict["global_feat"]["lr_init"]],
                                        dim=0).contiguous()
        keys = [k for k in self.params if k in extension_tensors and k != "global_feat"]
        if not keys:
            return
        src_of, n_keep = compact_plan(valid_mask)
        # group by extension length (one group in practice: every per-Gaussian tensor grows by the same rows)
        groups: dict[int, list] = {}
        for key in keys:
            ext = extension_tensors[key]
            groups.setdefault(0 if _is_empty(ext) else int(ext.shape[0]), []).append(key)
        for n_ext, gkeys in groups.items():
            jobs, slots = [], []
            for key in gkeys:
                param = self.params[key]
                ext = None if n_ext == 0 else extension_tensors[key]
                # the optimiser starts with N = 0 and shapeless placeholders: rows take their shape from the extension
                tail = tuple(ext.shape[1:]) if ext is not None else tuple(param["val"].shape[1:])
                jobs.append((param["val"].detach(), ext, 0, tail))
                slots.append((key, "val"))
                if key in _NO_MOMENTS:
                    continue
                jobs.append((param["exp_avg"], None, 0, tail))            # cat(m[mask], zeros_like(ext))
                slots.append((key, "exp_avg"))
                jobs.append((param["exp_avg_sq"], None, 0, tail))
                slots.append((key, "exp_avg_sq"))
                if key in self.lr_dict:                                   # cat(lr[mask], ones_like(ext) * lr_init)
                    jobs.append((param["lr"], None, self.lr_dict[key]["lr_init"], tail))
                    slots.append((key, "lr"))
            outs = compact_gather(src_of, n_keep, n_ext, jobs, n_mask=int(valid_mask.numel()))
            for (key, slot), out in zip(slots, outs):
                self.params[key][slot] = out
                if slot == "val" and key not in _NO_MOMENTS:
                    out.requires_grad = True
