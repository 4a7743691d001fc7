"""Drop-in for the reference's ``curope`` extension (croco/models/curope): ``rope_2d`` and the ``cuRoPE2D`` module."""
from artdeco_b200.mast3r.curope import cuRoPE2D, cuRoPE2D_func, rope_2d  # noqa: F401
