"""Drop-in for the reference's `freqencoder` package (freqencoder/__init__.py: `from .freq import FreqEncoder`)."""
from sanerf_hq_b200.encoders import FreqEncoder, freq_encode, _freq_encoder  # noqa: F401
