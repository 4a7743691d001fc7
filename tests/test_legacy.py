"""Legacy (Inria-convention) rasterizer surface — diff_gaussian_rasterization.GaussianRasterizer as the reference's web
viewer calls it (Reconstruct/webviewer/scene_models.py:559-605).
Oracle: oracle/legacy_torch.py (PARITY UNPINNED by the reference: the fork is not vendored).  Tolerance 1e-4 of scale for
images and gradients (outlier fraction 1e-3 on these tiny images for pixels whose skip/stop decision flips), ids exact."""
import math

import pytest
import torch

from artdeco_b200 import synthetic
from helpers import assert_close
from oracle import legacy_torch as lt


def _tiny(N=80, W=48, H=32, seed=4):
    sc = synthetic.raster_scene(N, seed=seed, z_range=(1.5, 6), extent=(2.5, 1.8), scale_range=(0.05, 0.4))
    V, K = synthetic.camera(W, H, view=5.0, focal=30.0)
    sc["opacities"][:5] = 0.002          # below 1/255: keeps a radius, emits nothing
    sc["means"][5:8, 2] = 0.1            # in front of the Inria near plane (0.2)
    return sc, V, K, W, H


def _settings(V, K, W, H, bg, sh_degree=3, scale_modifier=1.0):
    from artdeco_b200.legacy import GaussianRasterizationSettings
    tanx, tany = W / (2 * float(K[0, 0])), H / (2 * float(K[1, 1]))
    # projection-only matrix as the reference's call site builds it (getProjectionMatrix2, Reconstruct/utils.py:133-154,
    # passed transposed: webviewer/scene_models.py:549-566); the view matrix goes in per call, also transposed
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    P = torch.zeros(4, 4)
    zn, zf = 0.01, 100.0
    P[0, 0], P[1, 1] = 2 * fx / W, 2 * fy / H
    P[0, 2], P[1, 2] = (2 * cx - W) / W, (2 * cy - H) / H
    P[3, 2], P[2, 2], P[2, 3] = 1.0, zf / (zf - zn), -(zf * zn) / (zf - zn)
    campos = torch.inverse(V)[:3, 3]
    return GaussianRasterizationSettings(H, W, tanx, tany, bg, scale_modifier, P.t().contiguous(), sh_degree, campos,
                                         False, False)


def test_oracle_legacy_properties():
    sc, V, K, W, H = _tiny()
    bg = torch.tensor([0.2, 0.5, 0.8])
    empty = {k: sc[k][:0] for k in ("means", "quats", "scales", "opacities", "sh")}
    color, inv, main, radii = lt.rasterize(*empty.values(), V, K, W, H, bg)
    assert torch.allclose(color, bg.view(3, 1, 1).expand(3, H, W)) and (inv == 0).all() and (main == -1).all()
    # one big opaque splat straight ahead: alpha saturates at 0.99, inverse depth = 0.99 / z at its centre
    one = dict(means=torch.tensor([[0.0, 0.0, 4.0]]), quats=torch.tensor([[1.0, 0, 0, 0]]),
               scales=torch.full((1, 3), 2.0), opacities=torch.tensor([1.0]), sh=torch.zeros(1, 16, 3))
    V0, K0 = synthetic.camera(W, H, view=3.5, focal=30.0)
    color, inv, main, radii = lt.rasterize(*one.values(), V0, K0, W, H, bg)
    assert abs(float(inv[0, H // 2, W // 2]) - 0.99 / 4.0) < 2e-3 and int(main[0, H // 2, W // 2]) == 0
    assert int(radii[0]) == math.ceil(3 * math.sqrt((30.0 * 2.0 / 4.0) ** 2 + 0.3))


@pytest.mark.gpu
@pytest.mark.parametrize("scale_modifier", [1.0, 0.7])
def test_legacy_rasterizer_matches_oracle(cuda, scale_modifier):
    from artdeco_b200.legacy import GaussianRasterizer
    sc, V, K, W, H = _tiny()
    bg = torch.tensor([0.1, 0.3, 0.6])
    names = ("means", "quats", "scales", "opacities", "sh")
    ref_in = [sc[k].double().requires_grad_(True) for k in names]
    rc, ri, rm, rr = lt.rasterize(*ref_in, V.double(), K.double(), W, H, bg.double(), scale_modifier=scale_modifier)
    g = torch.Generator().manual_seed(9)
    vc, vi = torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g)
    ((rc * vc.double()).sum() + (ri * vi.double()).sum()).backward()

    t = {k: sc[k].to(cuda).requires_grad_(True) for k in names}
    rast = GaussianRasterizer(_settings(V, K, W, H, bg.to(cuda), scale_modifier=scale_modifier))
    m2d = torch.zeros(sc["means"].shape[0], 3, device=cuda, requires_grad=True)
    color, inv, main, radii = rast(t["means"], m2d, t["opacities"][:, None], t["sh"][:, :1], t["sh"][:, 1:],
                                   t["scales"], t["quats"], V.t().contiguous().to(cuda))
    assert color.shape == (3, H, W) and inv.shape == (1, H, W) and main.shape == (1, H, W) and main.dtype == torch.int32
    assert torch.equal(radii.cpu(), rr), "radii (ceil(3 sqrt(lambda_max)), 0 where culled)"
    assert (radii[:5] > 0).any() and (radii[5:8] == 0).all()
    assert_close(color, rc, what="color", max_outlier_frac=1e-3)
    assert_close(inv, ri, what="invdepth", max_outlier_frac=1e-3)
    assert (main.cpu() != rm).float().mean() <= 2e-3, "mainGaussID"
    ((color * vc.to(cuda)).sum() + (inv * vi.to(cuda)).sum()).backward()
    for k, r in zip(names, ref_in):
        assert_close(t[k].grad, r.grad, what=f"dL/d{k}", max_outlier_frac=1e-3)
    assert m2d.grad is not None and m2d.grad.shape == (sc["means"].shape[0], 3) and float(m2d.grad.abs().sum()) > 0


@pytest.mark.gpu
def test_legacy_shim_import_surface(cuda):
    import sys, pathlib
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1] / "shims"))
    try:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
        import diff_gaussian_rasterization as dgr
        assert callable(dgr.rasterize_gaussians) and callable(dgr.adamUpdate)
    finally:
        sys.path.pop(0)
    # 1080p smoke at scale: background where nothing lands, finite everywhere
    sc = synthetic.raster_scene(50000, seed=0)
    V, K = synthetic.camera(960, 540, view=2.0)
    rast = GaussianRasterizer(_settings(V, K, 960, 540, torch.zeros(3, device=cuda)))
    color, inv, main, radii = rast(sc["means"].to(cuda), None, sc["opacities"][:, None].to(cuda), sc["sh"][:, :1].to(cuda),
                                   sc["sh"][:, 1:].to(cuda), sc["scales"].to(cuda), sc["quats"].to(cuda),
                                   V.t().contiguous().to(cuda))
    assert torch.isfinite(color).all() and torch.isfinite(inv).all() and int((radii > 0).sum()) > 10000
    assert int(main.max()) < 50000 and int(main.min()) >= -1


def test_intrinsics_come_from_the_projection_only_matrix():
    """the legacy settings carry getProjectionMatrix2(...).T (projection only); the principal point must come
    back exactly for an off-centre cx/cy whatever the pose is (host-side logic, no kernel)."""
    from artdeco_b200.legacy import _intrinsics
    W, H = 640, 480
    K = torch.tensor([[500.0, 0, 300.0], [0, 480.0, 250.0], [0, 0, 1]])
    V, _ = synthetic.camera(W, H, view=7.0)          # non-identity pose: must not matter
    s = _settings(V, K, W, H, torch.zeros(3))
    Kb = _intrinsics(s, torch.device("cpu"))
    assert torch.allclose(Kb, K, atol=1e-4), Kb
