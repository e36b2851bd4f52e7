"""`shencoder.sphere_harmonics` module path of the reference."""
from sanerf_hq_b200.encoders import SHEncoder, sh_encode, _sh_encoder  # noqa: F401
