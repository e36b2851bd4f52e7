"""`freqencoder.freq` module path of the reference."""
from sanerf_hq_b200.encoders import FreqEncoder, freq_encode, _freq_encoder  # noqa: F401
