"""Drop-in for the reference's ``simple_knn`` package (Reconstruct/submodules/simple-knn)."""
from . import _C  # noqa: F401
