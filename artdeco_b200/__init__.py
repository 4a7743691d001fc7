"""artdeco_b200 — B200-native (sm_100a) kernels behind ARTDECO's hot-path operator surface.

Public operators (same names / argument meaning as the reference's, see each module's docstring):
  ssim.fused_ssim, ssim.FusedSSIMMap, ssim.fusedssim, ssim.fusedssim_backward      (fused_ssim)
  raster.rasterization                                 (gsplat.rendering.rasterization as ARTDECO calls it)
  adam.adamUpdate, adam.adamUpdateBasic                (diff_gaussian_rasterization, on-the-fly-nvs fork)
  knn.distCUDA2, knn.distIndex2, knn.distIndexQ        (simple_knn._C)
  cull.lod_select, cull.lod_cull                       (SceneModel.render's d_max cull)
  covmlp.cov_mlp_modulate                              (SceneModel.render's mlp_cov scale/rotation modulation)
  matching.iter_proj, matching.refine_matches, matching.match   (mast3r_slam_backends + utils_matching)
  optimizers.BaseAdam, optimizers.SparseGaussianAdam   (Reconstruct/scene/optimizers.py: fused step, one-launch add_and_prune)
  legacy.GaussianRasterizer, legacy.rasterize_gaussians  (diff_gaussian_rasterization legacy API, web viewer)
There is no CPU implementation: every operator raises unless a CUDA sm_100 device and the C-ABI library
``libartdeco_b200.so`` are available.
"""
from . import _lib  # noqa: F401
from . import adam, covmlp, cull, gn, knn, legacy, mast3r, matching, multiview, optimizers, parallel, peer, raster, scene, voxel  # noqa: F401  (each registers its C signatures)
from .adam import adamUpdate, adamUpdateBasic  # noqa: F401
from .covmlp import cov_mlp_modulate  # noqa: F401
from .cull import lod_cull, lod_select, weed_out_mask  # noqa: F401
from .scene import render_lod  # noqa: F401
from .voxel import update_voxel  # noqa: F401
from .knn import distCUDA2, distIndex2, distIndexQ  # noqa: F401
from .raster import rasterization  # noqa: F401
from .ssim import FusedSSIMMap, fused_ssim, fusedssim, fusedssim_backward  # noqa: F401

__version__ = "0.1.0"
