"""Drop-in for the reference's top-level `activation.py` (trunc_exp)."""
from sanerf_hq_b200.activation import trunc_exp, _trunc_exp  # noqa: F401
