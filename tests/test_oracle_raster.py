"""CPU: the C rasterizer oracle is pinned by torch.autograd on the pure-PyTorch restatement
(oracle/raster_torch.py), by the committed golden fixture, and by structural properties."""
import numpy as np
import pytest
import torch

import oracle
from oracle import raster_torch as rt
from artdeco_b200 import synthetic
from helpers import rel_err


def _tiny(N=60, W=32, H=32, seed=3):
    sc = synthetic.raster_scene(N, seed=seed, z_range=(2, 6), extent=(2.0, 2.0), scale_range=(0.05, 0.4))
    V, K = synthetic.camera(W, H, view=5.0, focal=30.0)
    sc["quats"] = sc["quats"] * 1.7  # un-normalised on purpose
    return sc, V, K, W, H


def test_c_oracle_matches_torch_autograd():
    sc, V, K, W, H = _tiny()
    means, quats, scales, opac, sh = [sc[k].double().requires_grad_(True)
                                      for k in ("means", "quats", "scales", "opacities", "sh")]
    Vd = V.double().requires_grad_(True)
    out, alpha, radii = rt.rasterization(means, quats, scales, opac, sh, Vd, K.double(), W, H)
    g = torch.Generator().manual_seed(5)
    vc = torch.randn(H, W, 4, generator=g, dtype=torch.float64)
    va = torch.randn(H, W, generator=g, dtype=torch.float64)
    ((out * vc).sum() + (alpha * va).sum()).backward()

    args = [sc[k].numpy() for k in ("means", "quats", "scales", "opacities", "sh")]
    f = oracle.rasterize_fwd(*args, V.numpy(), K.numpy(), W, H)
    assert (f["radii"] > 0).any(1).sum() > 20 and len(f["vals"]) > 50
    assert (f["radii"] == radii.numpy()).all()
    assert rel_err(f["colors"], out) < 2e-5 and rel_err(f["alphas"], alpha) < 2e-5
    b = oracle.rasterize_bwd(*args, V.numpy(), f, vc.float().numpy(), va.float().numpy())
    assert rel_err(b["v_means"], means.grad) < 5e-5
    assert rel_err(b["v_quats"], quats.grad) < 5e-5
    assert rel_err(b["v_scales"], scales.grad) < 5e-5
    assert rel_err(b["v_opac"], opac.grad) < 5e-5
    assert rel_err(b["v_sh"], sh.grad) < 5e-5
    # camera gradient: direct path + camera-centre path chained through inverse(viewmat) as autograd does
    Vt = V.double().requires_grad_(True)
    (torch.inverse(Vt)[:3, 3] * torch.tensor(b["v_campos"], dtype=torch.float64)).sum().backward()
    vV = b["v_viewmat"].astype(np.float64) + Vt.grad.numpy()
    assert rel_err(vV[:3], Vd.grad.numpy()[:3]) < 5e-5


def test_keys_sorted_and_offsets_consistent():
    sc = synthetic.raster_scene(5000, seed=1)
    V, K = synthetic.camera(640, 360, view=2.0)
    cam = oracle.make_cam(K.numpy(), 640, 360)
    radii, m2, d, con = oracle.project(sc["means"].numpy(), sc["quats"].numpy(), sc["scales"].numpy(),
                                       sc["opacities"].numpy(), V.numpy(), cam)
    tpg, keys, vals, offs = oracle.isect(radii, m2, d, 640, 360)
    assert tpg.sum() == len(keys) == len(vals) > 1000
    assert (np.diff(keys) >= 0).all()
    tiles = (keys >> 32)
    T = 40 * 23
    assert tiles.max() < T
    # offsets[t] is the first index with tile >= t
    assert (offs == np.searchsorted(tiles, np.arange(T))).all()
    # stable: equal keys keep emission (gaussian id) order
    same = np.nonzero(np.diff(keys) == 0)[0]
    assert (vals[same] < vals[same + 1]).all()
    # unsorted emission is Gaussian-major
    _, k2, v2, _ = oracle.isect(radii, m2, d, 640, 360, sort=False)
    assert (np.diff(v2) >= 0).all() and sorted(k2.tolist()) == keys.tolist()


def test_edge_cases_empty_and_culled():
    V, K = synthetic.camera(64, 48, focal=50.0)
    z = np.zeros
    f = oracle.rasterize_fwd(z((0, 3), np.float32), z((0, 4), np.float32), z((0, 3), np.float32), z(0, np.float32),
                             z((0, 16, 3), np.float32), V.numpy(), K.numpy(), 64, 48)
    assert f["colors"].shape == (48, 64, 4) and not f["colors"].any() and not f["alphas"].any()
    # behind the camera / transparent / far off-screen Gaussians are culled (radii == 0)
    means = np.array([[0, 0, -1.0], [0, 0, 5.0], [500.0, 0, 5.0], [0, 0, 5.0]], np.float32)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (4, 1))
    scales = np.full((4, 3), 0.1, np.float32)
    opac = np.array([0.9, 0.001, 0.9, 0.9], np.float32)
    f = oracle.rasterize_fwd(means, quats, scales, opac, z((4, 16, 3), np.float32), V.numpy(), K.numpy(), 64, 48)
    assert (f["radii"][:3] == 0).all() and (f["radii"][3] > 0).all()
    assert f["alphas"].max() > 0.5


def test_golden_fixture():
    """tests/golden/raster_small.npz was produced by tests/golden/make_raster_golden.py (C oracle, seed 7)."""
    import pathlib
    p = pathlib.Path(__file__).parent / "golden" / "raster_small.npz"
    g = np.load(p)
    sc = synthetic.raster_scene(int(g["N"]), seed=int(g["seed"]))
    V, K = synthetic.camera(int(g["W"]), int(g["H"]), view=float(g["view"]))
    f = oracle.rasterize_fwd(*[sc[k].numpy() for k in ("means", "quats", "scales", "opacities", "sh")], V.numpy(),
                             K.numpy(), int(g["W"]), int(g["H"]))
    assert (f["radii"] == g["radii"]).all()
    assert (f["keys"] == g["keys"]).all() and (f["vals"] == g["vals"]).all()
    assert (f["tile_offsets"] == g["tile_offsets"]).all()
    assert rel_err(f["colors"], g["colors"]) < 1e-6 and rel_err(f["alphas"], g["alphas"]) < 1e-6


def test_adam_oracle_matches_formula():
    rng = np.random.default_rng(0)
    p, g, m1, m2 = (rng.standard_normal((50, 3)).astype(np.float32) for _ in range(4))
    m2 = np.abs(m2)
    vis = rng.random(50) > 0.5
    lr = rng.random(50).astype(np.float32) * 1e-2
    P, M1, M2 = oracle.adam(p, g, m1, m2, vis, lr, 0.5, 0.99, 1e-15)
    e1 = 0.5 * m1 + 0.5 * g
    e2 = 0.99 * m2 + np.float32(0.01) * g * g
    ref = p - lr[:, None] * e1 / (np.sqrt(e2) + 1e-15)
    assert np.allclose(P[vis], ref[vis], rtol=1e-6) and (P[~vis] == p[~vis]).all()
    assert (M1[~vis] == m1[~vis]).all() and np.allclose(M2[vis], e2[vis], rtol=1e-6)
