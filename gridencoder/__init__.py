"""Drop-in for the reference's `gridencoder` package (gridencoder/__init__.py: `from .grid import GridEncoder`),
backed by hand-written sm_100a CUDA through the C ABI in include/sanerf_b200.h."""
from sanerf_hq_b200.encoders import GridEncoder, grid_encode, _grid_encode  # noqa: F401
