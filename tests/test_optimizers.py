"""SparseGaussianAdam mirror (Reconstruct/scene/optimizers.py): fused update+lr-decay and the one-launch add_and_prune
compaction against oracle/optimizers_ref.py.  Copies are exact; the Adam arithmetic is within 1e-6 relative (fp32)."""
import copy

import pytest
import torch

from helpers import assert_close
from oracle import optimizers_ref as oref

SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "scaling": (3,), "rotation": (4,), "opacity": (1,), "local_feat": (16,)}
LR_DICT = {"xyz": {"lr_init": 2e-4, "lr_decay": 0.997}, "f_rest": {"lr_init": 1e-3, "lr_decay": 0.99}}


def _fresh_params(dev):
    """As SceneModel builds them before the first key frame: N = 0 (h3dgsv3.py:150-200)."""
    params = {k: {"val": torch.empty((0,) + s, device=dev), "lr": 0.01} for k, s in SHAPES.items()}
    params["id"] = {"val": torch.empty(0, 1, dtype=torch.int64, device=dev), "lr": 0.0}
    params["d_max"] = {"val": torch.empty(0, 1, device=dev), "lr": 0.0}
    return params


def _ext(n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    ext = {k: torch.randn((n,) + s, generator=g).to(dev) for k, s in SHAPES.items()}
    ext["id"] = torch.arange(seed * 1000, seed * 1000 + n, dtype=torch.int64).view(n, 1).to(dev)
    ext["d_max"] = torch.rand(n, 1, generator=g).to(dev)
    return ext


def _clone_state(params):
    out = {}
    for k, pd in params.items():
        out[k] = {s: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for s, v in pd.items()}
    return out


def _assert_state_equal(a, b, exact=True):
    assert a.keys() == b.keys()
    for k in a:
        for s in ("val", "exp_avg", "exp_avg_sq", "lr"):
            if s not in b[k]:
                assert s not in a[k] or not isinstance(a[k][s], torch.Tensor) or s == "lr"
                continue
            x, y = a[k][s], b[k][s]
            if not isinstance(y, torch.Tensor):
                continue
            assert x.shape == y.shape and x.dtype == y.dtype, f"{k}.{s}: {x.shape} vs {y.shape}"
            if exact:
                assert torch.equal(x, y), f"{k}.{s}"
            else:
                assert_close(x, y, rtol=2e-6, what=f"{k}.{s}")


@pytest.mark.gpu
def test_sparse_gaussian_adam_matches_reference_semantics(cuda):
    from artdeco_b200.optimizers import SparseGaussianAdam
    ours = SparseGaussianAdam(_fresh_params(cuda), (0.5, 0.99), lr_dict=LR_DICT, device=cuda)
    ref = _fresh_params(cuda)
    # the oracle mirrors BaseAdam/SparseGaussianAdam.__init__ (optimizers.py:19-33,62-74)
    for k, pd in ref.items():
        pd["exp_avg"], pd["exp_avg_sq"] = torch.zeros_like(pd["val"]), torch.zeros_like(pd["val"])
        if k in ("id", "d_max"):
            continue
        pd["lr"] = torch.empty(0, device=cuda) if k in LR_DICT else torch.tensor(pd["lr"], dtype=torch.float, device=cuda)
    g = torch.Generator().manual_seed(0)
    N = 0
    for it, n_new in enumerate([5000, 0, 3000, 1, 20000]):
        # ---- add_and_prune: drop ~30 % of the rows, append n_new ----
        mask = (torch.rand(N, generator=g) > 0.3).to(cuda)
        ext = _ext(n_new, it + 1, cuda) if n_new else {k: torch.empty(0, device=cuda) for k in list(SHAPES) + ["id", "d_max"]}
        ours.add_and_prune(ext, mask)
        oref.add_and_prune(ref, LR_DICT, ext, mask)
        _assert_state_equal(ours.params, ref, exact=True)
        N = int(mask.sum()) + n_new
        assert ours.params["xyz"]["val"].shape == (N, 3) and ours.params["xyz"]["val"].requires_grad
        assert ours.params["f_rest"]["lr"].shape == (N, 15, 3) and not ours.params["id"]["val"].requires_grad
        # ---- two optimiser steps with a random visibility mask ----
        for s in range(2):
            vis = (torch.rand(N, generator=g) > 0.4).to(cuda)
            for k in SHAPES:
                gr = torch.randn(ours.params[k]["val"].shape, generator=g).to(cuda)
                ours.params[k]["val"].grad = gr.clone()
                ref[k]["val"].grad = gr.clone()
            ours.step(vis, N, None, 0)
            oref.step(ref, LR_DICT, (0.5, 0.99), 1e-15, vis)
            _assert_state_equal(ours.params, ref, exact=False)
        # re-synchronise so that rounding differences of the Adam arithmetic do not accumulate into the exact copy checks
        for k in ref:
            for s in ("val", "exp_avg", "exp_avg_sq", "lr"):
                if isinstance(ref[k].get(s), torch.Tensor) and isinstance(ours.params[k].get(s), torch.Tensor):
                    ref[k][s] = ours.params[k][s].detach().clone()
    # per-primitive learning rates decayed only on visible rows and never below 0.1 lr_init
    lr = ours.params["xyz"]["lr"]
    assert float(lr.min()) >= 0.1 * 2e-4 - 1e-12 and float(lr.max()) <= 2e-4 + 1e-12 and lr.unique().numel() > 2


@pytest.mark.gpu
def test_compaction_at_one_million_rows(cuda):
    """BASELINE scale: 1 M Gaussians, all 59 parameter floats + moments + lr + ids in one plan + one gather."""
    from artdeco_b200.optimizers import compact_gather, compact_plan
    N, n_ext = 1_000_000, 50_000
    g = torch.Generator().manual_seed(1)
    mask = (torch.rand(N, generator=g) > 0.2).to(cuda)
    src = {k: torch.randn((N,) + s, generator=g).to(cuda) for k, s in SHAPES.items()}
    ids = torch.arange(N, dtype=torch.int64, device=cuda).view(N, 1)
    ext = {k: torch.randn((n_ext,) + s, generator=g).to(cuda) for k, s in SHAPES.items()}
    src_of, n_keep = compact_plan(mask)
    assert n_keep == int(mask.sum())
    jobs = [(src[k], ext[k], 0, SHAPES[k]) for k in SHAPES] + [(src[k], None, 0.5, SHAPES[k]) for k in SHAPES] + [(ids, None, 0, (1,))]
    outs = compact_gather(src_of, n_keep, n_ext, jobs)
    for i, k in enumerate(SHAPES):
        assert torch.equal(outs[i], torch.cat([src[k][mask], ext[k]], 0)), k
        assert torch.equal(outs[len(SHAPES) + i][:n_keep], src[k][mask]) and bool((outs[len(SHAPES) + i][n_keep:] == 0.5).all())
    assert torch.equal(outs[-1][:n_keep, 0], torch.nonzero(mask)[:, 0]) and bool((outs[-1][n_keep:] == 0).all())


@pytest.mark.gpu
def test_base_adam_dense(cuda):
    from artdeco_b200.optimizers import BaseAdam
    g = torch.Generator().manual_seed(2)
    p = torch.randn(1000, 7, generator=g).to(cuda).requires_grad_(True)
    params = {"pose": {"val": p, "lr": 1e-3}}
    opt = BaseAdam(params, betas=(0.8, 0.99))
    ref_p, m, v = p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)
    for _ in range(3):
        gr = torch.randn(1000, 7, generator=g).to(cuda)
        p.grad = gr
        opt.step()
        oref.adam(ref_p, gr, m, v, torch.ones(1000, dtype=torch.bool, device=cuda), 1e-3, 0.8, 0.99, 1e-15)
    assert_close(p.detach(), ref_p, rtol=2e-6, what="BaseAdam")
    opt.zero_grad()
    assert p.grad is None
