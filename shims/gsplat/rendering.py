from artdeco_b200.raster import rasterization  # noqa: F401
