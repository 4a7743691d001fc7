"""artdeco_b200 — B200-native (sm_100a) kernels behind ARTDECO's hot-path operator surface.

Public operators (same names / argument meaning as the reference's, see each module's docstring):
  ssim.fused_ssim, ssim.FusedSSIMMap, ssim.fusedssim, ssim.fusedssim_backward
  rasterization.rasterization                      (gsplat.rendering.rasterization as ARTDECO calls it)
There is no CPU implementation: every operator raises unless a CUDA sm_100 device and the C-ABI library
``libartdeco_b200.so`` are available.
"""
from . import _lib  # noqa: F401
from . import raster  # noqa: F401  (registers C signatures)
from .raster import rasterization  # noqa: F401
from .ssim import FusedSSIMMap, fused_ssim, fusedssim, fusedssim_backward  # noqa: F401

__version__ = "0.1.0"
