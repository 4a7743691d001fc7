"""``simple_knn._C`` surface (ext.cpp:15-19): distCUDA2, distIndex2, distIndexQ."""
from artdeco_b200.knn import distCUDA2, distIndex2, distIndexQ  # noqa: F401
