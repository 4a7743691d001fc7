"""Drop-in for the reference's ``mast3r_slam_backends`` extension (VSLAM/backend/src/gn.cpp:84-122).

``iter_proj``, ``refine_matches``, ``gauss_newton_rays`` and ``gauss_newton_calib`` (everything ARTDECO calls:
VSLAM/utils_matching.py, VSLAM/mast3r_slam/global_opt.py:158,208) are served by artdeco_b200's kernels.  Any other attribute
(``gauss_newton_points``, which ARTDECO never calls) is DELEGATED to the reference's own compiled extension when one is
importable further down ``sys.path``; only when none exists does the lookup fail, loudly."""
import importlib.machinery
import importlib.util
import os
import sys

from artdeco_b200.gn import gauss_newton_calib, gauss_newton_rays  # noqa: F401
from artdeco_b200.matching import iter_proj, refine_matches  # noqa: F401

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_real = None


def _load_real():
    """Finds a ``mast3r_slam_backends`` module on sys.path that is NOT this shim and imports it under a private name."""
    global _real
    if _real is not None:
        return _real
    paths = [p for p in sys.path if os.path.abspath(p or ".") != _HERE]
    spec = importlib.machinery.PathFinder.find_spec("mast3r_slam_backends", paths)
    if spec is None or spec.loader is None:
        return None
    spec.name = "_mast3r_slam_backends_real"
    if hasattr(spec.loader, "name"):
        # extension modules export PyInit_mast3r_slam_backends: keep the loader's name so the init symbol resolves
        spec = importlib.util.spec_from_file_location("mast3r_slam_backends", spec.origin,
                                                      submodule_search_locations=spec.submodule_search_locations)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _real = mod
    return mod


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    real = _load_real()
    if real is not None and hasattr(real, name):
        return getattr(real, name)
    if name.startswith("gauss_newton"):
        raise NotImplementedError(f"mast3r_slam_backends.{name}: artdeco_b200 serves iter_proj / refine_matches / "
                                  "gauss_newton_rays / gauss_newton_calib; no compiled reference extension was found on "
                                  "sys.path to delegate this entry point to")
    raise AttributeError(name)
