"""Host-side mirror of ``simple_knn._C`` (Reconstruct/submodules/simple-knn/ext.cpp:15-19, spatial.cu:16-58):
``distCUDA2(points)``, ``distIndex2(points, K)``, ``distIndexQ(points, q_idx, n_idx, K)`` with the same argument
meaning and flat ``[P*K]`` return layout.  Differences, both allowed by the reference's own semantics: each row is
returned sorted by distance (the reference's slot order is traversal dependent, simple_knn.cu:391-421), and there
is no host synchronisation or allocation inside the call."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import i32, i64, vp

_lib.register("adb_knn_workspace_bytes", [i64, C.POINTER(C.c_size_t)])
_lib.register("adb_knn_mean3", [i64, vp, vp, vp, C.c_size_t, vp])
_lib.register("adb_knn_index", [i64, vp, i32, i64, vp, vp, vp, vp, vp, C.c_size_t, vp])


def _prep(points: torch.Tensor):
    _lib.require_cuda(points)
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must be [P, 3]")
    pts = points.detach().float().contiguous()
    nb = C.c_size_t(0)
    _lib.call("adb_knn_workspace_bytes", pts.shape[0], C.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=pts.device)
    return pts, ws


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance to the 3 nearest other points, [P] float32 (spatial.cu:16-26)."""
    pts, ws = _prep(points)
    P = pts.shape[0]
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.call("adb_knn_mean3", P, _lib.ptr(pts), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream())
    return out


def distIndex2(points: torch.Tensor, K: int):
    """Returns [dists[P*K] float32 (squared), ids[P*K] int32]; unfilled slots are FLT_MAX / -1 (spatial.cu:29-41)."""
    pts, ws = _prep(points)
    P = pts.shape[0]
    d = torch.empty(P * K, dtype=torch.float32, device=pts.device)
    ids = torch.empty(P * K, dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.call("adb_knn_index", P, _lib.ptr(pts), int(K), P, None, None, _lib.ptr(d), _lib.ptr(ids), _lib.ptr(ws),
                  ws.numel(), _lib.stream())
    return [d, ids]


def distIndexQ(points: torch.Tensor, q_indices: torch.Tensor, n_indices: torch.Tensor, K: int):
    """K nearest among ``points[n_indices]`` for each ``points[q_indices]`` (spatial.cu:44-58)."""
    pts, ws = _prep(points)
    P = pts.shape[0]
    q = q_indices.to(torch.int32).contiguous()
    cand = torch.zeros(P, dtype=torch.uint8, device=pts.device)
    cand[n_indices.long()] = 1
    Q = q.shape[0]
    d = torch.empty(Q * K, dtype=torch.float32, device=pts.device)
    ids = torch.empty(Q * K, dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.call("adb_knn_index", P, _lib.ptr(pts), int(K), Q, _lib.ptr(q), _lib.ptr(cand), _lib.ptr(d), _lib.ptr(ids),
                  _lib.ptr(ws), ws.numel(), _lib.stream())
    return [d, ids]
