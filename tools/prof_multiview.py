"""Minimal driver for ncu: a few EAGER steps of the bench's 8-view optimiser step (artdeco_b200.multiview.MultiViewStep),
no graph, no e2e, no CPU legs.  The kernel list of one step is what bench.py's value leg replays as a CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artdeco_b200 import synthetic  # noqa: E402
from artdeco_b200.multiview import MultiViewStep  # noqa: E402

N = int(os.environ.get("ADB_N", "1000000"))
steps = int(os.environ.get("ADB_STEPS", "2"))
views = int(os.environ.get("ADB_VIEWS", "8"))
W, H = 1920, 1080
dev = torch.device("cuda:0")
sc = synthetic.raster_scene(N, seed=0)
t = {k: sc[k].to(dev) for k in ("means", "quats", "scales", "opacities", "sh")}
cams = [synthetic.camera(W, H, view=float(v)) for v in range(views)]
eng = MultiViewStep(t, torch.stack([c[0] for c in cams]), torch.stack([c[1] for c in cams]), W, H, graph=False)
for j in range(views):
    vc, va = synthetic.upstream_grads(W, H, seed=1 + j)
    eng.v_colors[j].copy_(vc[0])
    eng.v_alphas[j].copy_(va[0, ..., 0])
eng.calibrate()
torch.cuda.synchronize()
if os.environ.get("ADB_CUPROF"):
    torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    eng.step()
torch.cuda.synchronize()
if os.environ.get("ADB_CUPROF"):
    torch.cuda.cudart().cudaProfilerStop()
print("n_isect", eng.check_overflow())
