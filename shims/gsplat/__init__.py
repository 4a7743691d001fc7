"""Drop-in for the one gsplat entry point ARTDECO calls (`gsplat.rendering.rasterization`, h3dgsv3.py:664)."""
from . import rendering  # noqa: F401
from .rendering import rasterization  # noqa: F401
