"""TEST INFRASTRUCTURE ONLY — the CPU oracle for artdeco_b200.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  Nothing under ``artdeco_b200/`` imports it; the product path has no CPU fallback.

Contents (each function cites the reference lines it restates in its own docstring / C header):
  raster_oracle.c   Gaussian-splat renderer fwd+bwd, sparse Adam           (gsplat / on-the-fly-nvs: un-vendored ->
                                                                            PARITY UNPINNED, see file header)
  knn_oracle.c      brute-force simple-knn results                          (pinned on GPU against oracle/_ref)
  raster_torch.py   tiny differentiable PyTorch renderer (pins the C oracle's gradients via autograd)
  ssim_ref.py       conv2d SSIM from the reference's own test (fused-ssim/tests/test.py:14-54) — PINNED:
                    that formulation is the known-answer check the reference asserts against.
  mast3r_ref.py     imports the reference MASt3R module (only where /root/reference exists) to generate goldens
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "liboracle.so"
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


class Cam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("W", C.c_int), ("H", C.c_int), ("eps2d", C.c_float), ("near_plane", C.c_float),
                ("far_plane", C.c_float), ("radius_clip", C.c_float)]


def build(force: bool = False) -> Path:
    srcs = [_DIR / "raster_oracle.c", _DIR / "knn_oracle.c", _DIR.parent / "include" / "adb_detmath.h"]
    if force or not _LIB.exists() or any(s.stat().st_mtime > _LIB.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_DIR), "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.adbo_isect_count.restype = C.c_int64
        _lib.adbo_num_threads.restype = C.c_int
        _lib.adbo_tile_bits.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().adbo_num_threads())


def set_num_threads(n: int) -> int:
    """Overrides OMP_NUM_THREADS for the C oracle (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().adbo_set_num_threads(C.c_int(int(n)))
    return num_threads()


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def make_cam(K, W, H, eps2d=0.01, near=0.01, far=1e10, radius_clip=0.0) -> Cam:
    K = np.asarray(K, dtype=np.float32)
    return Cam(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), int(W), int(H), eps2d, near, far,
               radius_clip)


def project(means, quats, scales, opac, viewmat, cam: Cam):
    means, quats, scales, opac, viewmat = map(_f, (means, quats, scales, opac, viewmat))
    N = means.shape[0]
    radii = np.zeros((N, 2), np.int32)
    means2d = np.zeros((N, 2), np.float32)
    depths = np.zeros(N, np.float32)
    conics = np.zeros((N, 3), np.float32)
    lib().adbo_project(C.c_int(N), means.ctypes, quats.ctypes, scales.ctypes, opac.ctypes, viewmat.ctypes,
                       C.byref(cam), radii.ctypes, means2d.ctypes, depths.ctypes, conics.ctypes)
    return radii, means2d, depths, conics


def campos_of(viewmat):
    """Camera centre = inverse(viewmat)[:3,3] (the live reference inverts the 4x4, h3dgsv3.py:627)."""
    return np.linalg.inv(np.asarray(viewmat, np.float64))[:3, 3].astype(np.float32)


def sh_fwd(means, campos, sh, radii, degree=3):
    means, campos, sh = map(_f, (means, campos, sh))
    N = means.shape[0]
    rgb = np.zeros((N, 3), np.float32)
    lib().adbo_sh_fwd(C.c_int(N), C.c_int(degree), means.ctypes, campos.ctypes, sh.ctypes,
                      np.ascontiguousarray(radii, np.int32).ctypes, rgb.ctypes)
    return rgb


def sh_bwd(means, campos, sh, radii, rgb, v_rgb, degree=3):
    means, campos, sh, rgb, v_rgb = map(_f, (means, campos, sh, rgb, v_rgb))
    N = means.shape[0]
    v_sh = np.zeros((N, 16, 3), np.float32)
    v_means = np.zeros((N, 3), np.float32)
    v_campos = np.zeros(3, np.float32)
    lib().adbo_sh_bwd(C.c_int(N), C.c_int(degree), means.ctypes, campos.ctypes, sh.ctypes,
                      np.ascontiguousarray(radii, np.int32).ctypes, rgb.ctypes, v_rgb.ctypes, v_sh.ctypes,
                      v_means.ctypes, v_campos.ctypes)
    return v_sh, v_means, v_campos


def isect(radii, means2d, depths, W, H, cam_id=0, n_cams=1, sort=True):
    """Returns (tiles_per_gauss[N], keys[I] int64, vals[I] int32, tile_offsets[T] int32)."""
    radii = np.ascontiguousarray(radii, np.int32)
    means2d, depths = _f(means2d), _f(depths)
    N = radii.shape[0]
    tpg = np.zeros(N, np.int32)
    total = int(lib().adbo_isect_count(C.c_int(N), radii.ctypes, means2d.ctypes, C.c_int(W), C.c_int(H), tpg.ctypes))
    keys = np.zeros(max(total, 1), np.int64)
    vals = np.zeros(max(total, 1), np.int32)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    offs = np.zeros(T, np.int32)
    lib().adbo_isect_sort(C.c_int(N), radii.ctypes, means2d.ctypes, depths.ctypes, C.c_int(W), C.c_int(H),
                          C.c_int(cam_id), C.c_int(n_cams), tpg.ctypes, C.c_int64(total), keys.ctypes, vals.ctypes,
                          offs.ctypes, C.c_int(1 if sort else 0))
    return tpg, keys[:total], vals[:total], offs


def blend_fwd(W, H, means2d, conics, opac, feats, vals, tile_offsets):
    means2d, conics, opac, feats = map(_f, (means2d, conics, opac, feats))
    vals = np.ascontiguousarray(vals, np.int32)
    CH = feats.shape[1]
    out = np.zeros((H, W, CH), np.float32)
    alphas = np.zeros((H, W), np.float32)
    last = np.zeros((H, W), np.int32)
    vv = vals if vals.size else np.zeros(1, np.int32)
    lib().adbo_blend_fwd(C.c_int(W), C.c_int(H), C.c_int(CH), means2d.ctypes, conics.ctypes, opac.ctypes,
                         feats.ctypes, vv.ctypes, C.c_int64(vals.size),
                         np.ascontiguousarray(tile_offsets, np.int32).ctypes, out.ctypes, alphas.ctypes, last.ctypes)
    return out, alphas, last


def blend_bwd(W, H, means2d, conics, opac, feats, vals, tile_offsets, alphas, last_ids, v_out, v_alphas):
    means2d, conics, opac, feats, alphas, v_out, v_alphas = map(_f, (means2d, conics, opac, feats, alphas, v_out, v_alphas))
    vals = np.ascontiguousarray(vals, np.int32)
    N, CH = feats.shape
    v_means2d = np.zeros((N, 2), np.float32)
    v_conics = np.zeros((N, 3), np.float32)
    v_opac = np.zeros(N, np.float32)
    v_feats = np.zeros((N, CH), np.float32)
    vv = vals if vals.size else np.zeros(1, np.int32)
    lib().adbo_blend_bwd(C.c_int(W), C.c_int(H), C.c_int(CH), C.c_int(N), means2d.ctypes, conics.ctypes, opac.ctypes,
                         feats.ctypes, vv.ctypes, C.c_int64(vals.size),
                         np.ascontiguousarray(tile_offsets, np.int32).ctypes, alphas.ctypes,
                         np.ascontiguousarray(last_ids, np.int32).ctypes, v_out.ctypes, v_alphas.ctypes,
                         v_means2d.ctypes, v_conics.ctypes, v_opac.ctypes, v_feats.ctypes)
    return v_means2d, v_conics, v_opac, v_feats


def project_bwd(means, quats, scales, opac, viewmat, cam: Cam, radii, v_means2d, v_depths, v_conics):
    means, quats, scales, opac, viewmat, v_means2d, v_depths, v_conics = map(
        _f, (means, quats, scales, opac, viewmat, v_means2d, v_depths, v_conics))
    N = means.shape[0]
    v_means = np.zeros((N, 3), np.float32)
    v_quats = np.zeros((N, 4), np.float32)
    v_scales = np.zeros((N, 3), np.float32)
    v_viewmat = np.zeros((4, 4), np.float32)
    lib().adbo_project_bwd(C.c_int(N), means.ctypes, quats.ctypes, scales.ctypes, opac.ctypes, viewmat.ctypes,
                           C.byref(cam), np.ascontiguousarray(radii, np.int32).ctypes, v_means2d.ctypes,
                           v_depths.ctypes, v_conics.ctypes, v_means.ctypes, v_quats.ctypes, v_scales.ctypes,
                           v_viewmat.ctypes)
    return v_means, v_quats, v_scales, v_viewmat


def rasterize_fwd(means, quats, scales, opac, sh, viewmat, K, W, H, sh_degree=3, eps2d=0.01):
    """Whole forward for one camera.  Returns a dict with every intermediate (all numpy)."""
    cam = make_cam(K, W, H, eps2d)
    radii, means2d, depths, conics = project(means, quats, scales, opac, viewmat, cam)
    campos = campos_of(viewmat)
    rgb = sh_fwd(means, campos, sh, radii, sh_degree)
    feats = np.concatenate([rgb, depths[:, None]], 1).astype(np.float32)
    tpg, keys, vals, offs = isect(radii, means2d, depths, W, H)
    out, alphas, last = blend_fwd(W, H, means2d, conics, opac, feats, vals, offs)
    return dict(cam=cam, radii=radii, means2d=means2d, depths=depths, conics=conics, campos=campos, rgb=rgb,
                feats=feats, tiles_per_gauss=tpg, keys=keys, vals=vals, tile_offsets=offs, colors=out,
                alphas=alphas, last_ids=last)


def rasterize_bwd(means, quats, scales, opac, sh, viewmat, fwd: dict, v_colors, v_alphas, sh_degree=3):
    """Whole backward.  Returns v_means, v_quats, v_scales, v_opac, v_sh, v_viewmat (4x4, excl. campos path),
    v_campos[3] (the caller chains it through inverse(viewmat) exactly as torch.autograd does in the reference)."""
    W, H = fwd["cam"].W, fwd["cam"].H
    v_m2, v_con, v_op, v_feats = blend_bwd(W, H, fwd["means2d"], fwd["conics"], opac, fwd["feats"], fwd["vals"],
                                           fwd["tile_offsets"], fwd["alphas"], fwd["last_ids"], v_colors, v_alphas)
    v_means, v_quats, v_scales, v_view = project_bwd(means, quats, scales, opac, viewmat, fwd["cam"], fwd["radii"],
                                                     v_m2, np.ascontiguousarray(v_feats[:, 3]), v_con)
    v_sh, v_means_sh, v_campos = sh_bwd(means, fwd["campos"], sh, fwd["radii"], fwd["rgb"],
                                        np.ascontiguousarray(v_feats[:, :3]), sh_degree)
    return dict(v_means=v_means + v_means_sh, v_quats=v_quats, v_scales=v_scales, v_opac=v_op, v_sh=v_sh,
                v_viewmat=v_view, v_campos=v_campos, v_means2d=v_m2, v_conics=v_con, v_feats=v_feats)


def adam(param, grad, m1, m2, visible, lr, b1, b2, eps):
    """In-place on copies; returns (param, m1, m2).  ``visible`` may be None."""
    param, grad, m1, m2 = (np.array(_f(a)) for a in (param, grad, m1, m2))
    N = param.shape[0]
    M = param.size // max(N, 1)
    lr = _f(np.atleast_1d(lr))
    vis = None if visible is None else np.ascontiguousarray(visible, np.uint8)
    lib().adbo_adam(C.c_int64(N), C.c_int64(M), param.ctypes, grad.ctypes, m1.ctypes, m2.ctypes,
                    vis.ctypes if vis is not None else None, lr.ctypes, C.c_int64(lr.size), C.c_float(b1),
                    C.c_float(b2), C.c_float(eps))
    return param, m1, m2


def knn_mean3(points):
    pts = _f(points)
    out = np.zeros(pts.shape[0], np.float32)
    lib().adbo_knn_mean3(C.c_int(pts.shape[0]), pts.ctypes, out.ctypes)
    return out


def knn_index(points, K, query_idx=None, candidate_mask=None):
    pts = _f(points)
    P = pts.shape[0]
    q = None if query_idx is None else np.ascontiguousarray(query_idx, np.int32)
    Q = P if q is None else q.shape[0]
    m = None if candidate_mask is None else np.ascontiguousarray(candidate_mask, np.uint8)
    d = np.zeros((Q, K), np.float32)
    ids = np.zeros((Q, K), np.int32)
    lib().adbo_knn_index(C.c_int(P), pts.ctypes, C.c_int(K), C.c_int(Q), q.ctypes if q is not None else None,
                         m.ctypes if m is not None else None, d.ctypes, ids.ctypes)
    return d, ids
