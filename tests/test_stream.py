"""BASELINE config 3 stand-in: the synthetic stream harness runs Frontend-like MASt3R + matching and the
mapper's render/optimise/densify loop on one GPU on separate streams.  Integration check only (finite numbers, growth,
every component called); the kernels themselves are pinned by their own parity tests."""
import pytest
import torch

from artdeco_b200 import synthetic
from oracle import mast3r_torch as mt


@pytest.mark.gpu
def test_synthetic_stream_runs_every_component(cuda):
    from artdeco_b200 import stream
    from artdeco_b200.mast3r import AsymmetricMASt3R
    cfg = mt.SMALL_CFG
    model = AsymmetricMASt3R(**cfg).load_state_dict(synthetic.det_weights(mt.param_shapes(cfg))).to(cuda)
    res = stream.run(cuda, frames=6, keyframe_every=2, grow=3000, mapper_iters=1, W=320, H=192, img=64, model=model)
    assert res["frames"] == 6 and res["n_keyframes"] == 3 and res["n_gaussians_final"] == 9000
    assert res["frontend_calls"] == 6 and res["mapper_calls"] == 6 and res["densify_calls"] == 3
    assert res["fps"] > 0 and all(res[k] > 0 for k in ("frontend_ms_mean", "mapper_ms_mean", "densify_ms_mean"))
