"""Drop-in for the reference's top-level `encoding.py` (get_encoder factory)."""
from sanerf_hq_b200.encoding import FreqEncoder_torch, get_encoder  # noqa: F401
