"""SparseGaussianAdam at 1 M Gaussians: our fused step / one-launch add_and_prune against the reference's op sequence
(oracle/optimizers_ref.py = Reconstruct/scene/optimizers.py restated in PyTorch; its adamUpdate replaced by OUR kernel so that
only the bookkeeping differs).  One JSON line; run under gpurun."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from artdeco_b200.adam import adamUpdate  # noqa: E402
from artdeco_b200.optimizers import SparseGaussianAdam  # noqa: E402
from oracle import optimizers_ref as oref  # noqa: E402
from test_optimizers import LR_DICT, SHAPES, _ext, _fresh_params  # noqa: E402


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    N, n_ext = 1_000_000, 50_000
    lr_dict = {k: {"lr_init": 1e-3, "lr_decay": 0.997} for k in SHAPES}      # every group on a per-primitive schedule
    opt = SparseGaussianAdam(_fresh_params(dev), (0.5, 0.99), lr_dict=lr_dict, device=dev)
    opt.add_and_prune(_ext(N, 1, dev), torch.zeros(0, dtype=torch.bool, device=dev))
    g = torch.Generator().manual_seed(0)
    vis = (torch.rand(N, generator=g) > 0.5).to(dev)
    for k in SHAPES:
        opt.params[k]["val"].grad = torch.randn(opt.params[k]["val"].shape, generator=g).to(dev)
    out = {"workload": f"SparseGaussianAdam, N={N}, 59+16 floats/Gaussian, 50 % visible; add_and_prune keeps 80 % and appends {n_ext}"}
    out["step_fused_ms"] = timed(lambda: opt.step(vis, N, None, 0))

    def ref_step():
        for k in SHAPES:
            pd = opt.params[k]
            p = pd["val"]
            adamUpdate(p, p.grad, pd["exp_avg"], pd["exp_avg_sq"], vis, pd["lr"], 0.5, 0.99, 1e-15, N, p.numel() // N)
            pd["lr"][vis] *= lr_dict[k]["lr_decay"]
            pd["lr"].clamp_min_(lr_dict[k]["lr_init"] * 0.1)
    with torch.no_grad():
        out["step_reference_ops_ms"] = timed(ref_step)
    floats = sum(int(torch.tensor(s).prod()) for s in SHAPES.values())
    out["step_algorithmic_GBps"] = 0.5 * N * floats * 4 * (5 + 4) / (out["step_fused_ms"] * 1e-3) / 1e9   # r: p,g,m,v,lr  w: p,m,v,lr

    mask = (torch.rand(N, generator=g) > 0.2).to(dev)
    ext = _ext(n_ext, 2, dev)
    state = {k: {s: v.detach().clone() for s, v in pd.items() if isinstance(v, torch.Tensor)} for k, pd in opt.params.items()}

    def restore(target):
        for k, pd in state.items():
            for s, v in pd.items():
                target[k][s] = v
    def ours():
        restore(opt.params)
        opt.add_and_prune(ext, mask)
    def theirs():
        restore(opt.params)
        oref.add_and_prune(opt.params, lr_dict, ext, mask)
    out["add_and_prune_ms"] = timed(ours, reps=5)
    out["add_and_prune_reference_ops_ms"] = timed(theirs, reps=5)
    words = floats * 4 + 2 + 1                      # params, m, v, lr + id (int64) + d_max
    out["add_and_prune_algorithmic_GBps"] = (0.8 * N + n_ext) * words * 4 * 2 / (out["add_and_prune_ms"] * 1e-3) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
