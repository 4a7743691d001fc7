"""Drop-in for the matching half of the reference's ``mast3r_slam_backends`` extension (VSLAM/backend/src/gn.cpp:84-112):
``iter_proj`` and ``refine_matches``.  The Gauss-Newton entry points (``gauss_newton_points / _rays / _calib``, SURVEY.md §8f
rank 4) are outside this build's scope and raise."""
from artdeco_b200.matching import iter_proj, refine_matches  # noqa: F401


def __getattr__(name):
    if name.startswith("gauss_newton"):
        raise NotImplementedError(f"mast3r_slam_backends.{name}: the global Gauss-Newton solver is out of scope "
                                  "(SURVEY.md §8f rank 4); only iter_proj / refine_matches are provided")
    raise AttributeError(name)
