"""Drop-in for the names the LIVE reference path imports from ``diff_gaussian_rasterization``
(Reconstruct/scene/optimizers.py:14): ``adamUpdate`` and ``adamUpdateBasic``.

The legacy Inria-convention rasterizer (``GaussianRasterizationSettings`` / ``GaussianRasterizer``), used only by
the off-path web viewer (Reconstruct/webviewer/scene_models.py:33-36, 559-605), is not provided
in this round: importing those names raises with a pointer to gsplat-style ``artdeco_b200.rasterization``."""
from artdeco_b200.adam import adamUpdate, adamUpdateBasic  # noqa: F401


def __getattr__(name):
    if name in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"):
        raise NotImplementedError(
            f"diff_gaussian_rasterization.{name}: the legacy Inria-convention rasterizer is outside this round's scope "
            "(SURVEY.md §8a R3, used only by the web viewer); the live path renders through gsplat.rendering.rasterization "
            "-> artdeco_b200.rasterization")
    raise AttributeError(name)
