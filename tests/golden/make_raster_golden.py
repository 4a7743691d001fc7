"""Generates tests/golden/raster_small.npz with the C oracle (run from the repo root:
``python tests/golden/make_raster_golden.py``).  The fixture pins the oracle against accidental change and gives
the GPU tests a committed, oracle-independent file to compare bit-exact integer outputs with."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from artdeco_b200 import synthetic  # noqa: E402

N, W, H, seed, view = 3000, 320, 192, 7, 1.0
sc = synthetic.raster_scene(N, seed=seed)
V, K = synthetic.camera(W, H, view=view)
f = oracle.rasterize_fwd(*[sc[k].numpy() for k in ("means", "quats", "scales", "opacities", "sh")], V.numpy(), K.numpy(), W, H)
np.savez_compressed(pathlib.Path(__file__).parent / "raster_small.npz", N=N, W=W, H=H, seed=seed, view=view,
                    radii=f["radii"], keys=f["keys"], vals=f["vals"], tile_offsets=f["tile_offsets"],
                    colors=f["colors"].astype(np.float16).astype(np.float32) * 0 + f["colors"], alphas=f["alphas"])
print("isect", len(f["keys"]), "visible", int((f["radii"] > 0).any(1).sum()))
