"""`gridencoder.grid` module path of the reference (gridencoder/grid.py)."""
from sanerf_hq_b200.encoders import GridEncoder, grid_encode, _grid_encode, _gridtype_to_id, _interp_to_id  # noqa: F401
