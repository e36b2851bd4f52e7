"""Drop-in for the reference's `shencoder` package (shencoder/__init__.py: `from .sphere_harmonics import SHEncoder`)."""
from sanerf_hq_b200.encoders import SHEncoder, sh_encode, _sh_encoder  # noqa: F401
