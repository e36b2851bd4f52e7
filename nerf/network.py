"""Drop-in for the reference's `nerf/network.py` (see nerf/renderer.py in this directory)."""
from sanerf_hq_b200.network import MLP, NeRFNetwork, SkipConnMLP  # noqa: F401
