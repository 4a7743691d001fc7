"""Drop-in for ``diff_gaussian_rasterization`` as the reference imports it:
  * ``adamUpdate`` / ``adamUpdateBasic`` — the LIVE path (Reconstruct/scene/optimizers.py:14,48-57,90-99,116-128,144-156);
  * ``GaussianRasterizationSettings`` / ``GaussianRasterizer`` / ``rasterize_gaussians`` — the legacy Inria-convention
    rasterizer the web viewer uses (Reconstruct/webviewer/scene_models.py:33-36, 559-605).
Everything runs through libartdeco_b200.so; there is no CPU fallback."""
from artdeco_b200.adam import adamUpdate, adamUpdateBasic  # noqa: F401
from artdeco_b200.legacy import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
